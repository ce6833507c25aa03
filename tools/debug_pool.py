import torch, torch.nn.functional as F
torch.manual_seed(0)
x0 = torch.randn(2, 3, 22, 25)
g0 = torch.randn(2, 3, 11, 13)
def run(x, dev, cl):
    x = x.to(dev)
    if cl:
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)   # NCHW view of NHWC memory
    x = x.detach().requires_grad_(True)
    y = F.avg_pool2d(x, kernel_size=2, padding=[0, 1])
    y.backward(g0.to(dev))
    return y.detach().cpu(), x.grad.detach().cpu()
yc, gc = run(x0, "cpu", False)
for cl in (False, True):
    y, g = run(x0, "cuda", cl)
    print(f"cuda channels_last-strided={cl}: fwd max diff vs cpu {(y-yc).abs().max().item():.2e}  bwd max diff vs cpu {(g-gc).abs().max().item():.2e}")
yl, gl = run(x0, "cpu", True)
print(f"cpu channels_last-strided: bwd diff {(gl-gc).abs().max().item():.2e}")
print(torch.__version__)
