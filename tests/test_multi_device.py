"""In-process multi-GPU (h2v_init(devices, n_dev)): handles hold one replica per device, the
host-facing batch entry points split their columns across the devices (column j -> device j mod G) and return results
in column order, `_dev` entry points run where their buffers live.  Needs >= 2 GPUs (skipped otherwise)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def two_devices(h2v):
    if h2v.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    h2v.init([0, 1])
    assert h2v.device_list() == [0, 1]
    yield h2v
    h2v.init(0)


def test_commit_batch_split_over_devices(two_devices):
    h = two_devices
    k, n = 10, 1 << 10
    bases = O.gen_bases(n)
    srs = h.ParamsKZG(k, None, bases)
    cols = [O.fr_fill(n, 300 + i, mode=i % 2) for i in range(11)]
    got = srs.commit_batch(cols)
    for g, c in zip(got, cols):
        assert (g == O.best_multiexp_affine(c, bases)).all()
    # single commits and short batches stay on the primary device
    assert (srs.commit_lagrange(cols[3]) == got[3]).all()
    srs.close()


def test_transform_batch_split_over_devices(two_devices):
    h = two_devices
    k = 9
    d, od = h.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    cols = [O.fr_fill(1 << k, 40 + i) for i in range(9)]
    for op, ref in ((h.OP_LAGRANGE_TO_COEFF, od.lagrange_to_coeff), (h.OP_COEFF_TO_EXTENDED, od.coeff_to_extended)):
        got = d.transform_batch(op, cols)
        for g, c in zip(got, cols):
            assert (g == ref(c)).all()
    d.close()


def test_dev_entry_points_run_where_the_buffers_live(two_devices):
    h = two_devices
    k, n = 9, 1 << 9
    bases = O.gen_bases(n)
    srs = h.ParamsKZG(k, None, bases)
    dom, od = h.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    cols = np.stack([O.fr_fill(n, 70 + i) for i in range(3)])
    for device in (0, 1):
        d_in = h.DeviceBuffer(cols.nbytes, device=device)
        d_in.upload(cols)
        d_out = h.DeviceBuffer(3 * 64, device=device)
        srs.commit_batch_dev(d_in.ptr, n, 3, n, d_out.ptr)
        got = d_out.download((3, 8))
        for i in range(3):
            assert (got[i] == O.best_multiexp_affine(cols[i], bases)).all(), (device, i)
        d_c = h.DeviceBuffer(cols.nbytes, device=device)
        dom.transform_dev(h.OP_LAGRANGE_TO_COEFF, d_in.ptr, n, d_c.ptr, n, 3)
        coef = d_c.download((3, n, 4))
        for i in range(3):
            assert (coef[i] == od.lagrange_to_coeff(cols[i])).all(), (device, i)
    srs.close()
    dom.close()


def test_init_rejects_bad_lists(h2v):
    with pytest.raises(ValueError):
        h2v.init([0, 0])
    with pytest.raises(ValueError):
        h2v.init([h2v.device_count()])
    h2v.init(0)
