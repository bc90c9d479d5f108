"""Builds csrc/*.cu into csrc/libgsr_b200.so for sm_100a with nvcc (in-tree, so the .so travels to the GPU box)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(CSRC, 'libgsr_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC'] + os.environ.get('GSR_NVCC_EXTRA', '').split()


def sources():
	return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
	if not os.path.exists(SO):
		return True
	deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
	return any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)


def build(force=False, verbose=False):
	"""One builder at a time (ranks of a multi-process launch all import the package: the first one in builds, the others wait on
	the lock and find the library fresh); objects go to a per-process directory and the library is renamed into place, so a
	reader never maps a half-written file."""
	if not force and not needs_build():
		return SO
	import fcntl
	with open(os.path.join(CSRC, '.build.lock'), 'w') as lock:
		fcntl.flock(lock, fcntl.LOCK_EX)
		try:
			if not force and not needs_build():
				return SO
			return _build_locked(verbose)
		finally:
			fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
	nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
	objs, procs = [], []
	os.makedirs(os.path.join(CSRC, 'build'), exist_ok=True)
	for src in sources():
		obj = os.path.join(CSRC, 'build', os.path.basename(src)[:-3] + '.o')
		objs.append(obj)
		cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
		procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
	for cmd, p in procs:
		out, _ = p.communicate()
		if verbose and out:
			print(out)
		if p.returncode != 0:
			raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + out)
	tmp = f'{SO}.{os.getpid()}.tmp'
	cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', tmp] + objs
	subprocess.run(cmd, check=True)
	os.replace(tmp, SO)
	return SO


if __name__ == '__main__':
	print(build(force=True, verbose=False))
