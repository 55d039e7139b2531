"""profiles/r2_traffic.json from an ncu CSV of `--metrics dram__bytes_read.sum,dram__bytes_write.sum` (one launch per
kernel of tools/profile_target.py <batch>).  Records the sha of the library that was profiled so that bench.py can say
whether its `roofline.traffic` comes from the build it is timing.

    python tools/ncu_traffic.py capture.csv <batch> [out.json]"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, batch = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, 'profiles', 'r2_traffic.json')
rows = [r for r in csv.reader(open(src, errors='replace')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
res = {}
for r in rows:
    if r is hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    name, metric, val, unit = d['Kernel Name'], d['Metric Name'], d['Metric Value'], d['Metric Unit']
    key = 'blind_rotate_kernel' if 'blind_rotate' in name else 'keyswitch_kernel' if 'keyswitch_kernel' in name else None
    if key is None or not metric.startswith('dram__bytes'):
        continue
    v = float(val.replace(',', ''))
    v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    res.setdefault(key, {})['dram_bytes_read' if 'read' in metric else 'dram_bytes_write'] = int(v)   # last launch wins
lib = os.environ.get('NUFHE_B200_LIB') or os.path.join(ROOT, 'nufhe_b200', 'csrc', 'libnufhe_b200.so')
try:
    import subprocess
    st = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'sass_stats.py'), '--json', '--lib', lib],
                                   capture_output=True, text=True, timeout=300).stdout.strip().splitlines()[-1])
    res['kernel_fingerprint'] = {'per_thread_step_total': st['per_thread_step_total'], 'phases': st['phases']}
except Exception:
    pass
res.update({'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, tools/profile_target.py %d (%s)' % (batch, os.path.basename(src)),
            'batch': batch, 'lib_sha': hashlib.sha256(open(lib, 'rb').read()).hexdigest()[:16]})
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res))
