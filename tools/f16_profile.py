"""Role counters of eval_mlp_f16_kernel (needs the triage build: ESTK_VARIANT=prof ESTK_EXTRA_FLAGS=-DESTK_TC_PROFILE
bash estorch_b200/csrc/build.sh, run with ESTK_LIBRARY=estorch_b200/lib/libestk_prof.so ESTK_TC_PROFILE=1)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend
from estorch_b200 import _capi
be = CudaBackend(torch.device("cuda", 0))
lib = _capi.load()
dims = [128, 512, 512, 512, 512, 288]
n = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
table = be.alloc(1 << 28); be.fill_noise_table(table, 42)
offs = be.alloc(pairs, dtype=torch.int64); order = be.alloc(pairs, dtype=torch.int32)
be.make_offsets(42, None, 0, 0, pairs, table.numel(), n, offs, order)
theta = torch.randn(n, device=be.device) * 0.05
obs, tgt = torch.randn(256, 128, device=be.device), torch.randn(256, 288, device=be.device)
ret = be.zeros(2 * pairs)
tb16 = be.alloc(table.numel(), dtype=torch.float16); assert be.shadow_f16(table, tb16) == 0
buf = (ctypes.c_ulonglong * 32)()
names = {0: "mma.total", 1: "mma.wait_h", 2: "mma.wait_full", 3: "mma.issue", 4: "prod.total", 7: "prod.wait_empty",
         8: "prod.form(+load waits)", 9: "prod.fence+arrive", 10: "epi.total", 11: "epi.obs", 12: "epi.bias+bar",
         13: "epi.wait_acc", 14: "epi.after_acc(move+tile1)", 15: "epi.tile0 drain+park"}
for rep in range(2):
    be.eval_mlp(dims, theta, table, offs, order, pairs, 0.02, obs, tgt, ret[:pairs], ret[pairs:], precision="f16", table16=tb16)
    lib.estk_debug_f16_profile(buf, 32)
tasks = -(-pairs * 2 // 74)
print(f"pairs={pairs} tasks/cluster~{tasks}")
for i, nm in names.items():
    print(f"  {nm:28s} {buf[i]:12d} cyc  {buf[i]/1.9e3/tasks:9.2f} us/task")
