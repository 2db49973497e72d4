import time, torch, sys
sys.path.insert(0, ".")
from chimeralm_b200.engine import Engine
from chimeralm_b200.weights import make_state_dict
sd = make_state_dict(0)
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
free0 = torch.cuda.mem_get_info()[0]
t0 = time.time(); eng = Engine(sd, device=0, max_batch=32, max_tokens=8193); torch.cuda.synchronize()
print("engine create + finalize + reserve: %.2f s" % (time.time() - t0))
print("device memory taken: %.0f MB" % ((free0 - torch.cuda.mem_get_info()[0]) / 1e6))
