import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend
from estorch_b200 import _capi
os.environ["ESTK_TC_DEBUG"] = os.environ.get("ESTK_TC_DEBUG", "8")
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
buf = (ctypes.c_ulonglong * 32)()
names = ["mma.total", "mma.wait_h", "mma.wait_full", "mma.issue", "prod.total", "prod.setup", "prod.stages", "prod.wait_empty",
         "prod.load+form+store", "prod.fence+arrive", "epi.total", "epi.obs", "epi.bias+bar+arrive", "epi.wait_acc", "epi.tmem+process", "epi.tmem_ld+wait"]
mode = os.environ.get("TC_MODE", "f16")
th16 = be.alloc(n, dtype=torch.bfloat16); be.shadow_bf16(theta, th16)
if mode == "f16":
    tb16 = be.alloc(table.numel(), dtype=torch.float16); assert be.shadow_f16(table, tb16) == 0
else:
    tb16 = be.alloc(table.numel(), dtype=torch.bfloat16); be.shadow_bf16(table, tb16)
kw = {"table16": tb16} if mode == "f16" else {"theta16": th16, "table16": tb16} if mode == "bf16s" else {}
for rep in range(2):
    be.eval_mlp(dims, theta, table, offs, order, pairs, 0.02, obs, tgt, ret[:pairs], ret[pairs:], precision=mode, **kw)
    lib.estk_debug_tc_profile(buf, 32)
tasks = -(-pairs * 2 // 74)
print(f"pairs={pairs} tasks/cluster~{tasks}")
for i, nm in enumerate(names):
    v = buf[i]
    print(f"  {nm:24s} {v:12d} cyc  {v/1.85e3/tasks:9.2f} us/task" if nm != "prod.stages" else f"  {nm:24s} {v:12d}")
