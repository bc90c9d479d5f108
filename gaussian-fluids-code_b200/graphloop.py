"""
Replay of a sync-free iteration body as a CUDA graph: the optimisation loops of the reference run hundreds to thousands of
iterations of a dozen short kernels each, and at the reference's sizes the host cannot enqueue them as fast as the GPU runs them.
GraphedLoop runs the body eagerly once per `unit` (which sizes every scratch buffer and warms the allocator), captures `unit`
calls into one graph and replays it; a remainder smaller than the unit runs eagerly.  The body must not synchronise with the host,
must take its random numbers from torch's CUDA generator or from device-resident counters (both advance correctly across replays)
and must read every scalar that changes between iterations from device memory (grid_scale, the Adam step count, the learning
rates: engine.FusedStepper keeps them in its state vector).
"""
import torch

from . import _lib

GRAPH_LAUNCHES = 0	# kernels of this library launched from graph replays (the library's host-side launch counter does not see them)


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
			side = torch.cuda.Stream()
			side.wait_stream(torch.cuda.current_stream())
			with torch.cuda.stream(side):
				l0 = lib.gsr_launch_count()
				for _ in range(self.unit):
					self.body()
				self.per_unit = lib.gsr_launch_count() - l0
			torch.cuda.current_stream().wait_stream(side)
			done += self.unit
			self.graph = torch.cuda.CUDAGraph()
			with torch.cuda.graph(self.graph):
				for _ in range(self.unit):
					self.body()
			GRAPH_LAUNCHES -= self.per_unit	# the capture pass bumped the host counter without running anything
		while done < n:
			if self.graph is not None and n - done >= self.unit:
				self.graph.replay()
				GRAPH_LAUNCHES += self.per_unit
				done += self.unit
			else:
				self.body()
				done += 1

	def release(self):
		self.graph = None
