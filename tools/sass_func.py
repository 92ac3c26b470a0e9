"""Print the SASS of one device function of the product library: python tools/sass_func.py <kernel> <function>"""
import re
import subprocess
import sys

kernel, func = sys.argv[1], sys.argv[2]
lib = 'hwang_b200/libhwang_b200.so'
elf = subprocess.run(['cuobjdump', '-elf', lib], capture_output=True, text=True).stdout
off = size = None
for line in elf.splitlines():
    m = re.match(r'\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|\d+)\s+(0x[0-9a-f]+|\d+)\s+0x(2|22|12)\s+\S+\s+\S+\s+(\S+)', line)
    if m and kernel in m.group(4) and m.group(4).startswith('$') and func in m.group(4).split('$')[-1]:
        off, size = int(m.group(1), 0), int(m.group(2), 0)
        break
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
inside = False
for line in sass.splitlines():
    if 'Function :' in line:
        inside = kernel in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        a = int(m.group(1), 16)
        if off <= a < off + size:
            print('%05x  %s' % (a - off, re.sub(r'\s+', ' ', m.group(2)).strip()))
