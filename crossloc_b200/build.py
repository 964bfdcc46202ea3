"""In-tree nvcc build of libcrossloc_b200.so (sm_100a only).

`python -m crossloc_b200.build` or `__graft_entry__.build()`.  The .so is written next to the package
(crossloc_b200/_C/libcrossloc_b200.so) so that it travels with the repository snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB = os.path.join(OUT_DIR, 'libcrossloc_b200.so')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = (["-DCL_DEBUG_TRAP"] if os.environ.get("CL_DEBUG_TRAP") else []) + ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-ccbin', '/usr/bin/g++']

# translation unit -> extra flags.  The pose solver is built without FMA contraction so that its double
# arithmetic rounds like the reference's x86 build of the same expressions.
UNITS = {
    'cabi.cu': [],
    'dsac.cu': ['--fmad=false'],
    'dsac_backward.cu': ['--fmad=false'],
    'cabi_cnn.cu': [],
    'net.cu': [],
    'conv_igemm.cu': [],
    'conv_wgrad_pf.cu': [],
    'cnn_pointwise.cu': [],
    'stem_tc.cu': [],
    'gn_backward.cu': [],
    'train_layout.cu': [],
    'frames.cu': [],
}


def _sources_digest():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(HERE, '..', 'include', 'crossloc_b200.h'), 'rb') as fh:
        h.update(fh.read())
    with open(__file__, 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, 'build.stamp')
    digest = _sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    objs = []
    procs = []
    for src, extra in UNITS.items():
        obj = os.path.join(OUT_DIR, src.replace('.cu', '.o'))
        cmd = [NVCC] + ARCH + COMMON + extra + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
            print(' '.join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(out)
        if p.returncode != 0:
            failed = True
            print('nvcc failed for %s' % src, file=sys.stderr)
    if failed:
        raise RuntimeError('crossloc_b200: nvcc build failed')
    cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart', '-lz', '-ccbin', '/usr/bin/g++']
    subprocess.check_call(cmd)
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
