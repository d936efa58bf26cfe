"""Where does stock PyTorch spend its time on the MiniGenerator forward at B = 65,536?  (one-off probe for bench.py's eager bar)"""
import time

import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
x = torch.randn(65536, 2, 16, device="cuda")
G = [torch.randn(*s, device="cuda") * 0.2 for s in [(4, 2, 3), (4,), (8, 4, 3), (8,), (4, 8, 3), (4,), (2, 4, 3), (2,)]]


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


e1 = F.leaky_relu(F.conv1d(x, G[0], G[1], stride=2, padding=1), 0.2)
b = F.leaky_relu(F.conv1d(e1, G[2], G[3], stride=2, padding=1), 0.2)
up1 = F.interpolate(b, scale_factor=2, mode="nearest")
d1 = F.leaky_relu(F.conv1d(up1, G[4], G[5], padding=1), 0.2) + e1
up2 = F.interpolate(d1, scale_factor=2, mode="nearest")
print("conv enc1   ", t(lambda: F.conv1d(x, G[0], G[1], stride=2, padding=1)))
print("lrelu       ", t(lambda: F.leaky_relu(e1, 0.2)))
print("conv bneck  ", t(lambda: F.conv1d(e1, G[2], G[3], stride=2, padding=1)))
print("interp 1    ", t(lambda: F.interpolate(b, scale_factor=2, mode="nearest")))
print("conv dec1   ", t(lambda: F.conv1d(up1, G[4], G[5], padding=1)))
print("interp 2    ", t(lambda: F.interpolate(d1, scale_factor=2, mode="nearest")))
print("conv out    ", t(lambda: F.conv1d(up2, G[6], G[7], padding=1)))
print("tanh        ", t(lambda: torch.tanh(up2)))
for B in (65535, 65536, 131072):
    xs = torch.randn(B, 4, 16, device="cuda")
    print("conv out at B =", B, t(lambda: F.conv1d(xs, G[6], G[7], padding=1)))
