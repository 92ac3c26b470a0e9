"""Memory-access instruction census of the product library's kernels (runs without a GPU):
  python tools/sass_vector_access.py > profiles/<round>_sass_vector_access.txt
Per kernel: how many global / generic loads and stores of each width the SASS holds, shared-memory and local-memory
(LDL/STL = register spills or thread-local arrays) accesses, and whether TMA / tensor-core instructions occur."""
import collections
import os
import re
import subprocess

lib = os.environ.get('HWB_PRODUCT_LIB', 'hwang_b200/libhwang_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern = None
census = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern.replace('(anonymous namespace)::', '')).split('::')[-1]
        census[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m or kern is None:
        continue
    op = m.group(1)
    census[kern]['instructions'] += 1
    base = op.split('.')[0]
    if base in ('LDG', 'STG', 'LD', 'ST', 'LDS', 'STS', 'LDL', 'STL', 'LDGSTS', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'ATOMG', 'ATOM', 'RED', 'LDC', 'LDSM'):
        width = '32'
        for w in ('128', '64', 'U16', 'S16', 'U8', 'S8'):
            if '.' + w in op:
                width = w.replace('U', '').replace('S', '')
                break
        census[kern]['%s.%s' % (base, width)] += 1
    if base.startswith(('HMMA', 'IMMA', 'UTCMMA', 'UTCHMMA', 'WGMMA', 'QMMA')):
        census[kern]['tensor-core'] += 1
print('library: %s' % lib)
for k, c in census.items():
    print('\n%s: %d instructions (%d bytes)' % (k, c['instructions'], 16 * c['instructions']))
    for name in sorted(n for n in c if n != 'instructions'):
        print('  %-12s %5d' % (name, c[name]))
    if not any(n.startswith(('LDL', 'STL')) for n in c):
        print('  (no local-memory access)')
