"""Builds the CUDA library in-tree:  python -m retargetvid_b200.build

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'rvb.cu')
OUT = os.path.join(HERE, 'lib', 'libretargetvid_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in ('rvb.cu', 'map_kernel.cuh', 'prim_kernel.cuh', 'fprim_kernel.cuh', 'prim_retire.inc', 'track_kernels.cuh', 'iou_kernel.cuh')]
DEPS.append(os.path.join(os.path.dirname(HERE), 'include', 'retargetvid_b200.h'))


def needs_build():
	if not os.path.isfile(OUT):
		return True
	t = os.path.getmtime(OUT)
	return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
	if not force and not needs_build():
		return OUT
	os.makedirs(os.path.dirname(OUT), exist_ok=True)
	nvcc = os.environ.get('NVCC', 'nvcc')
	cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
		'-Xcompiler', '-fPIC', '-shared', '-o', OUT, SRC]
	if verbose:
		cmd.insert(1, '-Xptxas')
		cmd.insert(2, '-v')
	r = subprocess.run(cmd, capture_output=True, text=True)
	if r.returncode != 0:
		sys.stderr.write(r.stdout + r.stderr)
		raise RuntimeError('nvcc failed building %s' % OUT)
	if verbose:
		sys.stderr.write(r.stderr)
	return OUT


if __name__ == '__main__':
	print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
