# interleaved A/B of two builds of the library on config 4: libhsc_b200_prev.so (the previous commit: `git stash; python -m
# hierarchical_sparse_coding_b200.build; cp .../libhsc_b200.so .../libhsc_b200_prev.so; git stash pop; rebuild`) against libhsc_b200.so,
# each with the window's first bulk copies issued late (default) or right after the pick (HSC_K2_EARLY_ISSUE=1)
show() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$1 ms/step %.2f value %.4g k1 %.2f k2 %.2f solo k2 %s clocks %s' % (d['ms_per_step'], d['value'], k['k1_ms'], k['k2_ms'], d['roofline'].get('ms_per_launch'), d['clocks']['sm_mhz']))"; }
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
  for lib in libhsc_b200_prev.so libhsc_b200.so; do
    for early in 0 1; do
      HSC_K2_EARLY_ISSUE=$early HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/$lib timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "$lib early=$early serial"
    done
  done
done
HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200.so timeout 300 python tools/latency_c1_c3.py 2>&1 | grep -E "c1_cmp|c3 hier"
HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prev.so timeout 300 python tools/latency_c1_c3.py 2>&1 | grep -E "c1_cmp|c3 hier"
