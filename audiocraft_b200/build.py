"""Build libaudiocraft_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m audiocraft_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libaudiocraft_b200.so')
STAMP = LIB + '.stamp'
SOURCES = ['api.cu', 'encodec.cu', 'lm.cu', 'lm_step.cu', 'lm_probe.cu']
HEADERS = ['common.cuh', 'gridbar.cuh', 'lm_step.cuh', os.path.join('..', '..', 'include', 'audiocraft_b200.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
         '-Xptxas', '-v']
if os.environ.get('ACB_STEP_NCW'):   # experiment: compute warps per CTA of the fused decode step (csrc/lm_step.cu)
    FLAGS.append('-DACB_STEP_NCW=' + os.environ['ACB_STEP_NCW'])
OBJ_SUFFIX = '.o'
if os.environ.get('ACB_BUILD_VARIANT'):   # experiment builds: extra -D flags (ACB_BUILD_DEFS) into libaudiocraft_b200_<variant>.so; run with ACB_LIB=<path>
    _v = os.environ['ACB_BUILD_VARIANT']
    FLAGS += os.environ.get('ACB_BUILD_DEFS', '').split()
    LIB = os.path.join(HERE, f'libaudiocraft_b200_{_v}.so')
    STAMP = LIB + '.stamp'
    OBJ_SUFFIX = f'.{_v}.o'
if os.environ.get('ACB_BUILD_TIMELINE') == '1':   # instrumented build: in-kernel %globaltimer stamps (see csrc/lm.cu tl_stamp)
    FLAGS.append('-DACB_TIMELINE')               # goes to its own file (never the product library): run with ACB_LIB=<that path>
    LIB = os.path.join(HERE, 'libaudiocraft_b200_timeline.so')
    STAMP = LIB + '.stamp'
    OBJ_SUFFIX = '.tl.o'


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    for f in SOURCES:
        obj = os.path.join(CSRC, f.replace('.cu', OBJ_SUFFIX))
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ['-c', os.path.join(CSRC, f), '-o', obj]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    logs = []
    for f, p in procs:
        out, _ = p.communicate()
        logs.append(f'== {f}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {f}:\n{out}')
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    with open(os.path.join(HERE, 'build.log' if OBJ_SUFFIX == '.o' else 'build_timeline.log'), 'w') as fh:
        fh.write('\n'.join(logs))
    if verbose:
        print('\n'.join(logs))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
