"""Copies the evidence set that tools/gpu_r2_profiles.sh left under gpurun_out/ into profiles/ (tracked) and refreshes
profiles/traffic.json (DRAM bytes per unit from the ncu --set full captures, read by bench.py for roofline.traffic)."""
import json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out'); P = os.path.join(ROOT, 'profiles')


def last_json_line(path):
    return json.loads([l for l in open(path).read().strip().splitlines() if l.startswith('{')][-1])


for src, dst in (('bench_r2.log', 'r2_bench_c4.json'), ('bench_r2_ref.log', 'r2_bench_c4_reference_arm.json'),
                 ('bench_r2_c2.log', 'r2_bench_c2.json'), ('bench_r2_c5.log', 'r2_bench_c5_ksvd.json')):
    json.dump(last_json_line(os.path.join(G, src)), open(os.path.join(P, dst), 'w'), indent=1)
for src, dst in (('prof_r2_k1.json', 'r2_k1_correlate_tc_fp16_ncu.json'), ('prof_r2_k2.json', 'r2_k2_pursuit_ncu.json'),
                 ('prof_r2_k2_c2.json', 'r2_k2_pursuit_c2_ncu.json'), ('prof_r2_k2_c5.json', 'r2_k2_pursuit_c5_ncu.json'),
                 ('prof_r2_k1_c5.json', 'r2_k1_correlate_tc_c5_ncu.json'), ('launches_r2.csv', 'r2_launches_c4_512signals.csv'),
                 ('pytest_gpu_r2.log', 'r2_gpu_tests_1gpu.log'), ('prof_r2_locomp.json', 'r2_locomp_fast_ncu.json'), ('locomp_mp_c4_c5.log', 'r2_locomp_vs_mp_c4_c5.txt'), ('prof_r2_k2_hot_lines.txt', 'r2_k2_hot_lines.txt')):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))


def dram_bytes(name):
    d = json.load(open(os.path.join(P, name)))['launches'][0]
    def gb(v):
        x, unit = v.split()
        return float(x) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[unit]
    return gb(d['dram__bytes_read.sum']) + gb(d['dram__bytes_write.sum'])


b4 = json.load(open(os.path.join(P, 'r2_bench_c4.json')))
b2 = json.load(open(os.path.join(P, 'r2_bench_c2.json')))
b5 = json.load(open(os.path.join(P, 'r2_bench_c5_ksvd.json')))
def atoms(b):
    return b['run']['selections_per_signal'] * b['config']['signals_per_gpu']
k1_4, k2_4 = dram_bytes('r2_k1_correlate_tc_fp16_ncu.json'), dram_bytes('r2_k2_pursuit_ncu.json')
k2_2, k2_5, k1_5 = dram_bytes('r2_k2_pursuit_c2_ncu.json'), dram_bytes('r2_k2_pursuit_c5_ncu.json'), dram_bytes('r2_k1_correlate_tc_c5_ncu.json')
traffic = {
    'c4': {'k1_dram_bytes_per_signal': k1_4 / 512, 'k2_dram_bytes_per_atom': k2_4 / atoms(b4),
           'source': 'profiles/r2_k1_correlate_tc_fp16_ncu.json (%.3f GB, 512 signals) and profiles/r2_k2_pursuit_ncu.json (%.3f GB, %d atoms): '
                     'dram__bytes_read.sum + dram__bytes_write.sum of one launch each, ncu --set full --clock-control none on '
                     '`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipeline 0` (tools/gpu_r2_profiles.sh)' % (k1_4 / 1e9, k2_4 / 1e9, atoms(b4))},
    'c2': {'k2_dram_bytes_per_atom': k2_2 / atoms(b2), 'k1_dram_bytes_per_signal': None,
           'source': 'profiles/r2_k2_pursuit_c2_ncu.json (%.3f GB, %d atoms)' % (k2_2 / 1e9, atoms(b2))},
    'c5': {'k2_dram_bytes_per_atom': k2_5 / atoms(b5), 'k1_dram_bytes_per_signal': k1_5 / b5['config']['signals_per_gpu'],
           'source': 'profiles/r2_k2_pursuit_c5_ncu.json (%.3f GB, %d atoms)' % (k2_5 / 1e9, atoms(b5))},
}
json.dump(traffic, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
print(json.dumps(traffic, indent=1))
