"""Builds libalona_b200.so (hand-written sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

    python -m alona_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libalona_b200.so')
SOURCES = ['ab_api.cu', 'ab_norm.cu', 'ab_hvg.cu', 'ab_irlb.cu', 'ab_knn.cu', 'ab_knn_exact.cu',
           'ab_knn_tc.cu', 'ab_snn.cu', 'ab_synth.cu', 'ab_io.cu', 'ab_qc.cu', 'ab_cluster.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--fmad=true', '-Xptxas', '-v']


def _digest():
    h = hashlib.sha256()
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(f.encode()); h.update(fh.read())
    with open(os.path.join(os.path.dirname(HERE), 'include', 'alona_b200.h'), 'rb') as fh:
        h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_digest(path=LIB):
    """The source digest the library at `path` was built from (embedded as the string 'AB_SOURCE_DIGEST=<hex>'),
    or None.  The digest travels inside the binary, so a stale .so next to newer sources is always detected - a
    side-car stamp file could not tell."""
    try:
        data = open(path, 'rb').read()
    except OSError:
        return None
    key = b'AB_SOURCE_DIGEST='
    i = data.find(key)
    return data[i + len(key):i + len(key) + 64].decode('ascii', 'replace') if i >= 0 else None


def is_fresh():
    return os.path.exists(LIB) and lib_digest() == _digest()


def build(force=False, verbose=False):
    """Compiles every .cu for sm_100a and links the shared library. Returns its path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and lib_digest() == dig:
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, s), '-o', obj]
        if s == 'ab_api.cu':
            cmd += ['-DAB_SOURCE_DIGEST="%s"' % dig]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        log.append('== %s\n%s' % (s, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % s)
        objs.append(obj)
    with open(os.path.join(objdir, 'ptxas.log'), 'w') as fh:
        fh.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart', '-ldl']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
