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
_CAPTURE_STREAM = None


import contextlib
import gc


@contextlib.contextmanager
def capture_guard():
	"""Around every stream capture: Python's cyclic collector must not run inside one.  A CUDAGraph (or any object owning CUDA
	resources) that sits in a reference cycle — a dead projector and the closures of its loop — is destroyed whenever the collector
	happens to run; if that is in the middle of ANOTHER capture, its clean-up is an operation the capturing stream does not permit
	and the capture is invalidated (seen in the full test suite: "operation not permitted when stream is capturing (function
	reset)").  So: collect first, then keep the collector off until the capture has ended."""
	gc.collect()
	was = gc.isenabled()
	gc.disable()
	try:
		yield
	finally:
		if was:
			gc.enable()


_SIDE_STREAM = None


def _side_stream():
	global _SIDE_STREAM
	if _SIDE_STREAM is None:
		_SIDE_STREAM = torch.cuda.Stream()
	return _SIDE_STREAM


class GraphedLoop:
	"""
	body()                      one iteration, or
	body(prepare())             when the iteration's inputs (sample batches) depend on nothing but the random stream: inside a
	                            captured unit the inputs of iteration j + 1 are then prepared on a second stream while iteration j
	                            computes.  Program order — and with it the order of the random draws — is unchanged, so the results
	                            are those of the plain loop.
	"""

	def __init__(self, body, unit=10, enabled=True, prepare=None):
		self.unit, self.enabled, self.prepare = unit, enabled, prepare
		self.body = body if prepare is None else (lambda: body(prepare()))
		self._body_in = body
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
		the allocator's cache every time turns the next allocations into cudaMalloc calls)"""
		global _CAPTURE_STREAM
		if _CAPTURE_STREAM is None:
			_CAPTURE_STREAM = torch.cuda.Stream()
		self.graph = torch.cuda.CUDAGraph()
		cur = torch.cuda.current_stream()
		_CAPTURE_STREAM.wait_stream(cur)
		with capture_guard(), torch.cuda.stream(_CAPTURE_STREAM):
			self.graph.capture_begin(capture_error_mode='thread_local')
			try:
				if self.prepare is None:
					for _ in range(self.unit):
						self.body()
				else:
					main, side = _CAPTURE_STREAM, _side_stream()
					keep = [self.prepare()]	# every prepared batch stays referenced until the capture ends: a block freed on one stream
					for j in range(self.unit):	# must not be handed out again while the other stream's kernels still read it
						if j + 1 < self.unit:
							side.wait_stream(main)	# (joins the capture; orders this prepare after the previous one's consumers were issued)
							with torch.cuda.stream(side):
								keep.append(self.prepare())
						self._body_in(keep[j])
						main.wait_stream(side)
					del keep
			finally:
				self.graph.capture_end()
		cur.wait_stream(_CAPTURE_STREAM)

	def release(self):
		"""drop the graph (and the closures: they and the objects they hold form a reference cycle with this loop)"""
		if self.graph is not None:
			torch.cuda.current_stream().synchronize()	# the last replay may still be running on the pool's memory
			self.graph = None
		self.enabled = False
		self.body = self._body_in = self.prepare = None
