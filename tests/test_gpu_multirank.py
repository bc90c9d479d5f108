"""
Two-GPU tests of the sharded paths, one process per GPU over NCCL — skipped on a one-GPU box (the CPU suite covers the
host logic with gloo, tests/test_multirank_cpu.py).  Run with:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one node')]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_case(case, timeout=600):
	with socket.socket() as s:
		s.bind(('127.0.0.1', 0))
		port = s.getsockname()[1]
	cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', str(port),
		   os.path.join(ROOT, 'tests', 'workers', 'two_rank_worker.py'), case]
	p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
	assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
	lines = [ln for ln in p.stdout.splitlines() if ln.startswith('RESULT ')]
	assert lines, p.stdout[-2000:]
	return json.loads(lines[-1][7:])


def test_sample_sharded_replicas_stay_bit_identical_and_p2p_equals_nccl():
	out = run_case('weak')
	assert out['nccl_replicas_bit_identical'] and out['p2p_replicas_bit_identical'] and out['p2p_vs_nccl_rel'] < 1e-4


def test_lattice_sharded_job_equals_the_single_gpu_job():
	out = run_case('strong')
	assert out['lattice_losses_rel'] < 1e-5


def test_lost_peer_times_out_with_zeros_and_an_error():
	assert run_case('timeout', timeout=300)['raised']
