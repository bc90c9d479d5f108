"""
Replay of a sync-free iteration body as a CUDA graph: the optimisation loops of the reference run hundreds to thousands of
iterations of a dozen short kernels each, and at the reference's sizes the host cannot enqueue them as fast as the GPU runs them.
GraphedLoop runs the body eagerly once per `unit` (which sizes every scratch buffer and warms the allocator), captures `unit`
calls into one graph and replays it; a remainder smaller than the unit runs eagerly.  The body must not synchronise with the host,
must take its random numbers from torch's CUDA generator or from device-resident counters (both advance correctly across replays)
and must read every scalar that changes between iterations from device memory (grid_scale, the Adam step count, the learning
rates: engine.FusedStepper keeps them in its state vector).
"""
import os
import time

import torch

from . import _lib

TRACE = bool(os.environ.get('GSR_GRAPHLOOP_TRACE'))	# development: print where a loop's wall time goes (synchronises)

GRAPH_LAUNCHES = 0	# kernels of this library launched from graph replays (the library's host-side launch counter does not see them)
_FREE_POOLS = []	# memory pools of released loops: the next capture reuses one instead of growing a new pool with cudaMalloc
_CAPTURE_STREAM = None


class GraphedLoop:
	def __init__(self, body, unit=10, enabled=True):
		self.body, self.unit, self.enabled = body, unit, enabled
		self.graph = None
		self.per_unit = 0

	def run(self, n):
		global GRAPH_LAUNCHES
		done = 0
		if self.enabled and self.graph is None and n >= 2 * self.unit:
			lib = _lib.lib()
			if TRACE:
				torch.cuda.synchronize(); t0 = time.perf_counter()
			side = torch.cuda.Stream()
			side.wait_stream(torch.cuda.current_stream())
			with torch.cuda.stream(side):
				l0 = lib.gsr_launch_count()
				for _ in range(self.unit):
					self.body()
				self.per_unit = lib.gsr_launch_count() - l0
			torch.cuda.current_stream().wait_stream(side)
			done += self.unit
			if TRACE:
				torch.cuda.synchronize(); t1 = time.perf_counter()
			self._capture()
			GRAPH_LAUNCHES -= self.per_unit	# the capture pass bumped the host counter without running anything
			if TRACE:
				torch.cuda.synchronize(); t2 = time.perf_counter()
				self.graph.replay(); torch.cuda.synchronize(); t3 = time.perf_counter()
				self.graph.replay(); torch.cuda.synchronize(); t4 = time.perf_counter()
				GRAPH_LAUNCHES += 2 * self.per_unit
				done += 2 * self.unit
				print(f'[graphloop] eager unit {1e3 * (t1 - t0):.1f} ms, capture {1e3 * (t2 - t1):.1f} ms, first replay {1e3 * (t3 - t2):.2f} ms, second {1e3 * (t4 - t3):.2f} ms ({self.unit} iterations)', flush=True)
		while done < n:
			if self.graph is not None and n - done >= self.unit:
				self.graph.replay()
				GRAPH_LAUNCHES += self.per_unit
				done += self.unit
			else:
				self.body()
				done += 1

	def _capture(self):
		"""torch.cuda.graph() without its synchronize + empty_cache prologue (an optimisation phase is captured once per frame: emptying
		the allocator's cache every time turns the next allocations into cudaMalloc calls), into a pool recycled from released loops"""
		global _CAPTURE_STREAM
		if _CAPTURE_STREAM is None:
			_CAPTURE_STREAM = torch.cuda.Stream()
		self.pool = _FREE_POOLS.pop() if _FREE_POOLS else torch.cuda.graph_pool_handle()
		self.graph = torch.cuda.CUDAGraph()
		cur = torch.cuda.current_stream()
		_CAPTURE_STREAM.wait_stream(cur)
		with torch.cuda.stream(_CAPTURE_STREAM):
			self.graph.capture_begin(pool=self.pool, capture_error_mode='thread_local')
			try:
				for _ in range(self.unit):
					self.body()
			finally:
				self.graph.capture_end()
		cur.wait_stream(_CAPTURE_STREAM)

	def release(self):
		if self.graph is not None:
			torch.cuda.current_stream().synchronize()	# the last replay may still be running on the pool's memory
			self.graph = None
			_FREE_POOLS.append(self.pool)
