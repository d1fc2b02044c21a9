"""SASS opcode histogram of the shipped library (cuobjdump -sass): the evidence that the tensor-core
kernels use tcgen05 (UTCHMMA), TMEM loads (LDTM), bulk async copies (UBLKCP) and mbarriers (SYNCS).

  python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'nerfies_b200', 'libnerfies_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
  m = re.match(r'\s*Function : (\S+)', line)
  if m:
    kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
    hist[kern] = collections.Counter()
    continue
  m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
  if m and kern:
    hist[kern][m.group(1)] += 1
KEY = ('UTCHMMA', 'UTCBAR', 'LDTM', 'UBLKCP', 'UTMALDG', 'STTM', 'SYNCS', 'FFMA', 'DFMA', 'HMMA', 'F2FP', 'STS', 'LDS', 'MUFU')
print(f'# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)')
print('# columns: total instructions | ' + ' '.join(KEY) + '  (prefix match; e.g. UTCHMMA counts UTCHMMA and UTCHMMA.2CTA)')
for k, h in hist.items():
  tot = sum(h.values())
  cols = [sum(v for op, v in h.items() if op.startswith(p)) for p in KEY]
  print(f'{k[:70]:70s} {tot:7d} | ' + ' '.join(f'{c:5d}' for c in cols))
print('\n# full opcode list of the tensor-core kernels')
for k, h in hist.items():
  if 'field_tc_kernel<1, 0>' in k or 'field_x3_kernel' in k or k.endswith('field_tc_kernel<1>'):
    print(f'## {k}')
    print('   ' + ', '.join(f'{op} {n}' for op, n in sorted(h.items(), key=lambda t: -t[1])))
