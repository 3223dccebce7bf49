"""Build libgcrl_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with
the repo snapshot to the GPU box).

    python -m graph_colouring_with_rl_b200.build [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.environ.get('GCRL_LIB_OUT') or os.path.join(PKG_DIR, 'libgcrl_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC,-fvisibility=hidden',
          '--expt-relaxed-constexpr', '-Xptxas', '-v'] + os.environ.get('GCRL_NVCC_FLAGS', '').split()
# GCRL_NVCC_FLAGS: extra flags for diagnostic builds, e.g. -DGCRL_WS_TRACE_BUILD (pipeline timeline of CTA 0 of the
# warp-specialised kernels, dumped when GCRL_WS_TRACE=<file> is set) or -DGCRL_MBAR_TIMEOUT (trap on a stuck mbarrier)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        glob.glob(os.path.join(os.path.dirname(PKG_DIR), 'include', '*.h')) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu into one shared library.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    objdir = os.path.join(PKG_DIR, 'build' if not os.environ.get('GCRL_LIB_OUT') else 'build_diag')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + ARCH_FLAGS + CFLAGS + ['-c', src, '-o', obj]
        procs.append((src, obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs = []
    log = []
    for src, obj, cmd, p in procs:
        out = p.communicate()[0].decode()
        log.append('$ ' + ' '.join(cmd) + '\n' + out)
        if p.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError('nvcc failed on %s' % src)
        objs.append(obj)
    link = [NVCC] + ARCH_FLAGS + ['-shared', '-o', LIB_PATH] + objs + ['-lcudart']
    out = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    log.append('$ ' + ' '.join(link) + '\n' + out.stdout.decode())
    if out.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError('link failed')
    with open(os.path.join(objdir, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
