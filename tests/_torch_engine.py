"""A plain-torch stand-in for ``mae_clip_b200.dist.CudaStripEngine`` (TEST INFRASTRUCTURE).

Restates the three row-strip sweeps of the C ABI (``mc_clip_stats / mc_clip_rowloss / mc_clip_bwd``,
include/mae_clip_b200.h) so the collective choreography of ``mae_clip_b200/dist.py`` can be
exercised on CPU with gloo at world_size 2.  Never imported by the product."""
import torch


class TorchStripEngine:
    def __init__(self, dtype=torch.float64):
        self.dtype = dtype

    def _strips(self, I_all, T_all, b, row_offset, tau):
        I = I_all.to(self.dtype)
        T = T_all.to(self.dtype)
        Il, Tl = I[row_offset:row_offset + b], T[row_offset:row_offset + b]
        S = Tl @ I.T / tau            # owned logits rows
        St = Il @ T.T / tau           # St[i, j] = S[j, gi]
        Z = (Il @ I.T + Tl @ T.T) * (tau / 2)
        return S, St, Z

    def stats(self, I_all, T_all, b, row_offset, tau):
        S, St, Z = self._strips(I_all, T_all, b, row_offset, tau)
        out = torch.stack([torch.logsumexp(S, 1), torch.logsumexp(St, 1), torch.logsumexp(Z, 1)])
        return out.float(), None  # (this stand-in keeps S, so it does not need the sum_j P_ij S_ij vector)

    def rowloss(self, I_all, T_all, planes, b, row_offset, tau, stats_all):
        S, St, Z = self._strips(I_all, T_all, b, row_offset, tau)
        B = I_all.shape[0]
        r, c, rz = (v.to(self.dtype) for v in stats_all)
        ri, rzi = r[row_offset:row_offset + b], rz[row_offset:row_offset + b]
        P = torch.exp(Z - rzi[:, None])
        G = -(2 * S - ri[:, None] - c[None, :]) / (2 * B)
        g = (P * G).sum(1)
        q = torch.exp(Z - rz[None, :]).sum(1)  # column sums of P through the symmetry of Z
        return torch.stack([g, q]).float(), g.sum().reshape(1).float()

    def bwd(self, I_all, T_all, planes, b, row_offset, tau, stats_all, gq_all, grad_loss):
        S, St, Z = self._strips(I_all, T_all, b, row_offset, tau)
        B = I_all.shape[0]
        I, T = I_all.to(self.dtype), T_all.to(self.dtype)
        r, c, rz = (v.to(self.dtype) for v in stats_all)
        g, q = (v.to(self.dtype) for v in gq_all)
        sl = slice(row_offset, row_offset + b)
        P, Pt = torch.exp(Z - rz[sl, None]), torch.exp(Z - rz[None, :])
        dS = (torch.exp(S - r[sl, None]) + torch.exp(S - c[None, :]) * q[None, :] - 2 * P) / (2 * B)
        dSt = (torch.exp(St - r[None, :]) + torch.exp(St - c[sl, None]) * q[sl, None] - 2 * Pt) / (2 * B)
        G = -(2 * S - r[sl, None] - c[None, :]) / (2 * B)
        Gt = -(2 * St - r[None, :] - c[sl, None]) / (2 * B)
        dZs = P * (G - g[sl, None]) + Pt * (Gt - g[None, :])
        gl = grad_loss.to(self.dtype)
        dT = gl * (dS @ I / tau + (tau / 2) * dZs @ T)
        dI = gl * (dSt @ T / tau + (tau / 2) * dZs @ I)
        return dI.float(), dT.float()
