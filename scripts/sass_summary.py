"""SASS evidence per kernel of libgcrl_b200.so (cuobjdump -sass): counts of the Blackwell mnemonics that prove the
tcgen05 / TMEM / TMA path (B200_PROFILING.md): UTCHMMA (tcgen05.mma kind::f16), UTMALDG / UTMASTG / UTMAREDG (TMA
tensor load / store / reduce), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus
registers and spills from the build log.    python scripts/sass_summary.py > profiles/r02_sass_summary.json"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'graph_colouring_with_rl_b200', 'libgcrl_b200.so')
MNEMONICS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA', 'RED', 'ATOM',
             'BAR', 'LDS', 'STS', 'LDG', 'STG']


def clean(name):
    name = name.replace('(bool)', '').replace('(int)', '')
    return re.sub(r'\(.*', '', name).replace('void gcrl::', '').replace('gcrl::', '')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    names = sorted(set(re.findall(r'Function : (\S+)', sass)))
    if names:
        out = subprocess.run(['cu++filt'] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            cur = clean(cur)
            kernels[cur] = {'instructions': 0}
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]['instructions'] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + '.'):
                kernels[cur][mn] = kernels[cur].get(mn, 0) + 1
    # registers / spills from the build log (-Xptxas -v)
    log = os.path.join(ROOT, 'graph_colouring_with_rl_b200', 'build', 'build.log')
    if os.path.isfile(log):
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", text, re.S):
            nm = subprocess.run(['cu++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            nm = clean(nm)
            if nm in kernels:
                kernels[nm].update(registers=int(m.group(5)), stack_bytes=int(m.group(2)), spill_store_bytes=int(m.group(3)),
                                   spill_load_bytes=int(m.group(4)))
    tot = {mn: sum(k.get(mn, 0) for k in kernels.values()) for mn in MNEMONICS[:9]}
    json.dump({'library': os.path.relpath(LIB, ROOT), 'how': 'cuobjdump -sass + cu++filt; registers/spills from nvcc -Xptxas -v',
               'totals': tot, 'kernels': {k: v for k, v in sorted(kernels.items())}}, sys.stdout, indent=1)


if __name__ == '__main__':
    main()
