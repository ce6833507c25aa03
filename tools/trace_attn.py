"""Per-event clock64 trace of one attention CTA (needs a -DTCL_ATTN_TRACE build: make EXTRA=-DTCL_ATTN_TRACE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tclight_b200 import ops
from tclight_b200 import _lib
lib = _lib.load_tuning_lib()      # needs a -DTCL_ATTN_TRACE tuning build: make -C tclight_b200/csrc tuning EXTRA=-DTCL_ATTN_TRACE
ops.lib = lib

var = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda"); dt = torch.bfloat16
B, H, T, d = 2, 8, 47520, 40
dp = ops.head_pad(d); Tp = (T + 7) // 8 * 8
q = torch.randn(B, H, Tp, dp, device=dev).to(dt); k = torch.randn(B, H, Tp, dp, device=dev).to(dt)
vt = torch.randn(B, H, dp, Tp, device=dev).to(dt)
buf = torch.zeros(192, dtype=torch.int64, device=dev)
lib.tcl_debug_attention_variant(var)
ops.attention(q, k, vt, T, T, d)
lib.tcl_debug_attention_trace(buf.data_ptr())
ops.attention(q, k, vt, T, T, d)
torch.cuda.synchronize()
t = buf.cpu().view(3, 8, 8)
t0 = int(t[0, 0, 0])
names = {0: ["waitS", "gotS", "ldS", "expdone", "gotPe", "Pstored"], 1: ["waitS", "gotS", "ldS", "expdone", "gotPe", "Pstored"],
         2: ["q0 waitSe", "q0 gotSe", "q0 waitPf", "q0 gotPf", "q1 waitSe", "q1 gotSe", "q1 waitPf", "q1 gotPf"]}
print("variant", var)
for j in range(8):
    for role in range(3):
        ev = [(names[role][e], int(t[role, j, e]) - t0) for e in range(len(names[role]))]
        print(f"j={64+j} role={role}: " + "  ".join(f"{n}={v}" for n, v in ev))
