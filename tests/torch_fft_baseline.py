"""Test/measurement helper (not product): kymatio's Scattering2D dataflow on the GPU with
torch.fft (cuFFT) and elementwise torch ops, batched over signals — what moving kymatio's torch frontend to
CUDA would execute: every intermediate is a separate kernel with an HBM round trip.  Filters come from the
oracle's bank.  Used as a second, independent fp32 cross-check of the fused kernel and as the 'library path'
timing next to it (python tests/torch_fft_baseline.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import Scattering2D as OracleScattering2D  # noqa: E402


class TorchFFTScattering2D:
    def __init__(self, J, shape, L=8, max_order=2, device="cuda"):
        o = OracleScattering2D(J=J, shape=shape, L=L, max_order=max_order, cache_filters=True)
        self.J, self.L, self.max_order, self.o = J, L, max_order, o
        self.M, self.N = shape
        self.Mp, self.Np = o._M_padded, o._N_padded
        dev = torch.device(device)
        self.phi = [torch.from_numpy(l).to(dev) for l in o.phi["levels"]]
        self.psi = [dict(j=p["j"], levels=[torch.from_numpy(l).to(dev) for l in p["levels"]]) for p in o.psi]
        self.pad = ((self.Mp - self.M) // 2, (self.Mp - self.M + 1) // 2, (self.Np - self.N) // 2, (self.Np - self.N + 1) // 2)

    @staticmethod
    def _fold(x, k):
        if k == 1:
            return x
        s = x.shape
        return x.reshape(s[:-2] + (k, s[-2] // k, k, s[-1] // k)).mean(dim=(-4, -2))

    def __call__(self, x):
        """x: [B, M, N] float32 CUDA tensor -> [B, K, h, w]."""
        J = self.J
        xp = torch.nn.functional.pad(x[:, None], (self.pad[2], self.pad[3], self.pad[0], self.pad[1]), mode="reflect")[:, 0]
        U0 = torch.fft.fft2(xp)
        out0 = [torch.fft.ifft2(self._fold(U0 * self.phi[0], 2 ** J)).real[..., 1:-1, 1:-1]]
        out1, out2 = [], []
        for p1 in self.psi:
            j1 = p1["j"]
            U1 = torch.fft.fft2(torch.fft.ifft2(self._fold(U0 * p1["levels"][0], 2 ** j1)).abs())
            out1.append(torch.fft.ifft2(self._fold(U1 * self.phi[j1], 2 ** (J - j1))).real[..., 1:-1, 1:-1])
            if self.max_order < 2:
                continue
            for p2 in self.psi:
                j2 = p2["j"]
                if j2 <= j1:
                    continue
                U2 = torch.fft.fft2(torch.fft.ifft2(self._fold(U1 * p2["levels"][j1], 2 ** (j2 - j1))).abs())
                out2.append(torch.fft.ifft2(self._fold(U2 * self.phi[j2], 2 ** (J - j2))).real[..., 1:-1, 1:-1])
        return torch.stack(out0 + out1 + out2, dim=1)


if __name__ == "__main__":
    import time
    import wst_b200
    for M, J, B in [(32, 2, 4096), (64, 3, 2048), (128, 4, 512)]:
        x = torch.rand(B, 3, M, M, device="cuda")
        S = TorchFFTScattering2D(J, (M, M))
        ref = S(x[:8].reshape(-1, M, M))
        plan = wst_b200.get_plan(M, M, J, 8)
        _, maps = plan.forward(x[:8].contiguous(), False, True)
        err = float((maps.reshape(ref.shape) - ref).abs().max() / ref.abs().max())
        for fn, name in ((lambda: S(x.reshape(-1, M, M)).mean(dim=(-2, -1)), "torch.fft (cuFFT + elementwise)"),
                         (lambda: plan.forward(x), "wst_b200 fused cascade")):
            fn(); torch.cuda.synchronize(); t = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
            print(f"{M}x{M} J={J}: {name:34s} {B / dt:12.0f} patches/s   (max diff vs torch.fft path {err:.1e})")
