import ctypes as C, json, sys
sys.path.insert(0, '.')
import torch
from gaussian_fluids_code_b200 import _lib
lib = _lib.lib()
torch.zeros(1, device='cuda')
out = {}
for name in ('gsr_peak_fma', 'gsr_peak_mufu'):
	best = 0.
	for it in range(5):
		v = C.c_double(0.)
		rc = getattr(lib, name)(C.c_int(20000), C.byref(v), _lib.stream())
		assert rc == 0, rc
		best = max(best, v.value)
	out[name] = best
p = torch.cuda.get_device_properties(0)
out['sms'] = p.multi_processor_count
out['name'] = p.name
print(json.dumps(out))
