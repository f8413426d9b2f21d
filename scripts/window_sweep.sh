for w in 5 3 1; do
  H2V_RED_WEIGHT=$w python bench.py --no-extras 2>/dev/null > gpurun_out/w$w.json
  python -c "
import json;d=json.load(open('gpurun_out/w$w.json'));print('w',$w,d['value'],d['e2e']['value'],d['roofline']['window_bits'],d['roofline']['windows'])"
done
