"""DRAM traffic per launch of the decode kernels, from an `ncu --set full` capture (run here, no GPU needed):

  python tools/kernel_traffic.py gpurun_out/x.ncu-rep <frames of the captured workload> > profiles/kernel_traffic.json

bench.py reads `roofline.traffic` for its dominant kernel from that file (only when the frame count matches)."""
import csv, json, subprocess, sys

rep, frames = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
out = {}
for r in rows[2:]:
    m = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = m['Kernel Name'].split('::')[-1].split('(')[0]
    try:
        rd = float(m['dram__bytes_read.sum'].replace(',', '')) * scale.get(u['dram__bytes_read.sum'], 1)
        wr = float(m['dram__bytes_write.sum'].replace(',', '')) * scale.get(u['dram__bytes_write.sum'], 1)
    except ValueError:
        continue
    k = out.setdefault(name, {'launches': 0, 'dram_bytes_read': 0.0, 'dram_bytes_write': 0.0})
    k['launches'] += 1; k['dram_bytes_read'] += rd; k['dram_bytes_write'] += wr
for k in out.values():
    k['dram_bytes_per_launch'] = (k['dram_bytes_read'] + k['dram_bytes_write']) / k['launches']
print(json.dumps({'frames': frames, 'source': rep, 'kernels': out}, indent=1))
