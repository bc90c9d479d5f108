"""FP32 / LDS pipe probes (development tool): what the SM sustains for the operand shapes of the evaluation kernels"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import _lib
lib = _lib.lib()
torch.zeros(1, device='cuda')
names = ['ffma_3reg', 'ffma_2reg_same', 'ffma_reg_uniform_reg', 'fadd', 'fmul', 'candidate_test_x4', 'lds128_uniform_x3', 'ffma2_3pairs(scalar fma/s)', 'ffma2_pair_bcast_pair(scalar fma/s)', 'candidate_test_x4_f32x2']
out = {}
for w, nm in enumerate(names):
	best = 0.
	for _ in range(3):
		v = C.c_double(0.)
		rc = lib.gsr_pipe_probe(C.c_int(w), C.c_int(20000), C.byref(v), _lib.stream())
		assert rc == 0, rc
		best = max(best, v.value)
	out[nm] = best / 1e12	# T thread-ops / s
p = torch.cuda.get_device_properties(0)
slots = p.multi_processor_count * 128 * 1.965e9 / 1e12
out['issue_slots_T_per_s_at_1965MHz'] = slots
print(json.dumps(out))
