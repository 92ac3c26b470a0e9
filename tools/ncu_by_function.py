"""Attribute the per-instruction counters of an ncu report to the device functions of a kernel.

  ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
  python tools/ncu_by_function.py sass.csv entropy_cabac_ip_kernel [units] [--hot N]

The SASS page lists the kernel and its (noinline) callees by absolute address; the ELF symbol table of the product library
gives every callee's offset and size inside the kernel's .text section, so offset = address - first address.
`units` (e.g. the number of macroblocks the launch decoded) turns the counts into a per-unit figure.
The report must have been taken from the library as it is built now.
"""
import csv
import os
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith('--')]
hot = int(sys.argv[sys.argv.index('--hot') + 1]) if '--hot' in sys.argv else 0
if '--hot' in sys.argv:
    args = [a for a in args if a != str(hot)] if str(hot) in args[2:] else args
path, kernel = args[0], args[1]
units = float(args[2]) if len(args) > 2 else 0.0

elf = subprocess.run(['cuobjdump', '-elf', os.environ.get('HWB_PRODUCT_LIB', 'hwang_b200/libhwang_b200.so')], capture_output=True, text=True).stdout
funcs = []  # (offset, size, name)
ksize = 0
for line in elf.splitlines():
    m = re.match(r'\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|\d+)\s+(0x[0-9a-f]+|\d+)\s+0x(2|22|12)\s+\S+\s+\S+\s+(\S+)', line)
    if not m or kernel not in m.group(4):
        continue
    off, size, name = int(m.group(1), 0), int(m.group(2), 0), m.group(4)
    if name.startswith('$'):
        short = subprocess.run(['c++filt', name.split('$')[-1]], capture_output=True, text=True).stdout.strip()
        short = re.sub(r'\(.*', '', short).split('::')[-1]
        funcs.append((off, size, short))
    else:
        ksize = size
funcs.sort()

rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[hdr]
ia, isrc, iex, ismp = H.index('Address'), H.index('Source'), H.index('Instructions Executed'), H.index('# Samples')
inst = []
for r in rows[hdr + 1:]:
    if len(r) != len(H):
        continue
    inst.append((int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)))
base = inst[0][0]


def owner(off):
    for o, s, n in funcs:
        if o <= off < o + s:
            return n
    return '(kernel body + inlined)'


tot_ex = sum(i[2] for i in inst)
tot_smp = sum(i[3] for i in inst)
agg = {}
for a, src, ex, smp in inst:
    f = owner(a - base)
    e = agg.setdefault(f, [0, 0, 0, 0])
    e[0] += ex; e[1] += smp; e[2] += 1
    if units and ex >= 0.25 * units:
        e[3] += 16  # bytes of code executed at least once per four units: what the instruction caches have to hold
print('%-28s %8s %7s %7s %9s %s' % ('function', 'instr', 'exec%', 'smp%', 'cyc/inst', 'exec/unit  hot bytes' if units else ''))
for f, (ex, smp, n, hb) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print('%-28s %8d %7.2f %7.2f %9.2f %s' % (f, n, 100.0 * ex / tot_ex, 100.0 * smp / max(1, tot_smp),
                                              (smp / max(1, tot_smp)) / max(1e-12, ex / tot_ex), ('%9.1f %6d' % (ex / units, hb)) if units else ''))
if units:
    print('hot code (executed >= 0.25 times per unit): %d bytes' % sum(v[3] for v in agg.values()))
print('total warp instructions executed: %d%s' % (tot_ex, (' = %.1f per unit' % (tot_ex / units)) if units else ''))
if hot:
    print('\nhottest instructions:')
    for a, src, ex, smp in sorted(inst, key=lambda x: -x[3])[:hot]:
        print('%06x %-26s ex %10d smp %6d  %s' % (a - base, owner(a - base), ex, smp, src[:80]))
