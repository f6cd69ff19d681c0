"""Device-side synthetic workloads for the benches (setup only, never timed).

eeg_like_distance_matrices follows SURVEY.md §8(d) config (a)/(b): per recording
x = A(47x8)/sqrt(8) @ S(8xT) + 0.5*E, windows of 250 samples, Pearson correlation,
d = sqrt(2(1-r)) in float64, cast to float32 (the arithmetic of
/root/reference/notebooks/2_graph_construction.ipynb:86-122)."""
import torch


def eeg_like_distance_matrices(B, n=47, win=250, k=8, noise=0.5, seed=20261018, device="cuda", chunk=8192):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((B, n, n), dtype=torch.float32, device=device)
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        A = torch.randn((nb, n, k), generator=g, device=device, dtype=torch.float64) / k ** 0.5
        S = torch.randn((nb, k, win), generator=g, device=device, dtype=torch.float64)
        E = torch.randn((nb, n, win), generator=g, device=device, dtype=torch.float64)
        x = A @ S + noise * E
        x = x - x.mean(dim=2, keepdim=True)
        c = x @ x.transpose(1, 2) / (win - 1)
        sd = torch.sqrt(torch.diagonal(c, dim1=1, dim2=2))
        r = (c / sd[:, :, None] / sd[:, None, :]).clamp_(-1, 1)
        d = torch.sqrt(2 * (1 - r)).clamp_min_(0)
        d.diagonal(dim1=1, dim2=2).zero_()
        out[b0:b0 + nb] = d.float()
    return out
