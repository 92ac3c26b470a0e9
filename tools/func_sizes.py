"""Print the SASS size of every device function of the product library (code-footprint work: the entropy kernel is
instruction-fetch bound).  Usage: python tools/func_sizes.py [substring]"""
import re
import subprocess
import sys

out = subprocess.run(['cuobjdump', '-elf', __import__('os').environ.get('HWB_PRODUCT_LIB', 'hwang_b200/libhwang_b200.so')], capture_output=True, text=True).stdout
pat = sys.argv[1] if len(sys.argv) > 1 else ''
rows = []
for line in out.splitlines():
    m = re.match(r'\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|\d+)\s+(0x[0-9a-f]+|\d+)\s+0x(2|22|12)\s+\S+\s+\S+\s+(\S+)', line)
    if not m:
        continue
    size = int(m.group(2), 0)
    name = m.group(4)
    if pat not in name:
        continue
    short = subprocess.run(['c++filt', name.split('$')[-1] if '$' in name else name], capture_output=True, text=True).stdout.strip()
    kern = re.search(r'cbf86f86\d+([a-z0-9_]+_kernel)', name.split('$')[1] if name.startswith('$') else name)
    rows.append((size, (kern.group(1) if kern else '?'), short[:90]))
tot = {}
for s, k, n in sorted(rows, key=lambda r: (r[1], -r[0])):
    print('%7d  %-26s %s' % (s, k, n))
