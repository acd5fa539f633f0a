import ctypes, sys
sys.path.insert(0, "/root/repo/vit-is-all-you-need_b200")
import torch
from b200vit import ops, _cabi
from b200vit._cabi import ptr
lib = _cabi.ensure_device(0)
B, N, H = 256, 197, 12
d = H * 64
qkv = torch.randn(B, N, 3 * d, device="cuda").to(torch.bfloat16)
do = torch.randn(B, N, d, device="cuda").to(torch.bfloat16)
o, lse = ops.flash_attn_fwd(qkv, B, N, H, False)
dqkv = torch.empty(B, N, 3 * d, device="cuda", dtype=torch.bfloat16)
ws = torch.zeros(B * N * d, device="cuda", dtype=torch.float32)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    rc = lib.b200vit_flash_attn_bwd(ptr(qkv), ptr(o), ptr(do), ptr(lse), ptr(dqkv), B, N, H, 0, 0, ptr(ws), ws.numel() * 4, st)
    assert rc == 0
torch.cuda.synchronize()
t = ws[:256].view(torch.int64).cpu().tolist()
base = min(x for x in t if x > 0)
names_m = {0: "full", 1: "c0 S/dP issued", 2: "c0 pds0", 3: "c0 pds1", 4: "c0 dV/dK issued", 5: "c0 dkv_free(prev)", 9: "c1 S/dP issued", 10: "c1 pds0", 11: "c1 pds1", 12: "c1 dkv issued", 13: "c1 dkv_free", 20: "dq issued"}
print("MMA thread:")
for k in sorted(names_m, key=lambda k: t[k]):
    print(f"  {t[k]-base:7d}  {names_m[k]}")
for wg in (0, 1):
    print(f"WG{wg} (warp {4+4*wg} lane 0):")
    names_c = {0: "unit start", 1: "stats done", 2: "c0 sdp_full", 3: "c0 P/dS done", 4: "c0 dkv_full", 5: "c0 readout done", 10: "c1 sdp_full", 11: "c1 P/dS done", 12: "c1 dkv_full", 13: "c1 readout done", 20: "dq_full", 21: "dq readout done"}
    for k in sorted(names_c, key=lambda k: t[16 + wg * 16 + k] if False else t[(32 + wg * 32 + k)]):
        print(f"  {t[32 + wg*32 + k]-base:7d}  {names_c[k]}")
