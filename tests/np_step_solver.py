"""Test-only NumPy (float64) implementation of the step-solver protocol of background-subtraction_b200/dist.py.
It lets the sharded driver's collective choreography run on CPU ranks under gloo; the arithmetic of one shard is the
oracle's (this file lives under tests/ and is never imported by the product)."""
import numpy as np
import torch

from oracle import alm_oracle as O


class NumpyStepSolver:
    def __init__(self, D_local, rows, cols_local, n, m_global, delta=10, mu_scale=12.5, rho=1.6, tol=1e-7, max_iter=500,
                 labels=None, lambdas_by_frame=None, graph_cols=None):
        """labels (int [n][m_local], 0 = complement) + lambdas_by_frame select the group-sparse solver
        (/root/reference/group_sparse_RPCA.py:45-126: l2 blocks, mu0 = 1.25/||D||_2, break on rank 0)."""
        self.D = np.asfortranarray(D_local, dtype=np.float64)          # m_local x n
        self.rows, self.cols, self.n = rows, cols_local, n
        self.groups = O.flat_groups_nonoverlap((rows, cols_local), (3, 3))
        self.lam = 1.0 / (np.sqrt(max(m_global, n)) * delta)
        self.d = min(m_global, n)
        self.mu_scale, self.rho, self.tol, self.max_iter = mu_scale, rho, tol, max_iter
        self.sum_buf = torch.zeros(n * n + 16, dtype=torch.float64)
        self.max_buf = torch.zeros(8, dtype=torch.float64)
        self.ngram = n * n
        self.log = []
        self.labels, self.lams = labels, lambdas_by_frame
        self.graph_cols = graph_cols            # overlapping 3x3 windows of the WHOLE frame (rows x graph_cols): prox on whole frames
        if graph_cols is not None:
            self.gc_full = O.graph_all_groups((rows, graph_cols), (3, 3))
        self.bsums = None
        if labels is not None:
            self.mu_scale = 1.25
            self.nlab = 1 + max(len(l) for l in lambdas_by_frame)
            self.bsums = torch.zeros(n * self.nlab, dtype=torch.float64)
            self.nbl = 100.0 * self.lam

    def block_sums_view(self): return self.bsums

    def graph_split(self): return (self.rows, self.graph_cols) if self.graph_cols is not None else None

    def prox_buffers(self):
        return torch.from_numpy(self.G_S.T), torch.from_numpy(self.S.T)          # [n][m_local] views of the F-order matrices

    def frames_like(self, nf, ldf): return torch.zeros((max(nf, 1), ldf), dtype=torch.float64)

    def prox_frames(self, Uf, Vf, rows, cols, nf):
        if self.done_flag or nf == 0: return
        m = rows * cols
        V = O.prox_graph(np.asfortranarray(Uf[:nf, :m].numpy().T), self.lam / self.mu, self.gc_full, tol=1e-13)
        Vf[:nf, :m] = torch.from_numpy(np.ascontiguousarray(V.T))

    # -- views used by the driver's all-reduces
    def gram_view(self): return self.sum_buf[:self.ngram]
    def tail_view(self): return self.sum_buf[self.ngram:self.ngram + 4]
    def sum_view(self): return self.sum_buf[:self.ngram + 9]
    def mask_tail_view(self): return self.sum_buf[self.ngram + 4:self.ngram + 8]
    def max_view(self): return self.max_buf

    def init_local(self):
        self.sum_buf.zero_(); self.max_buf.zero_()
        self.sum_buf[:self.ngram] = torch.from_numpy((self.D.T @ self.D).ravel())
        self.max_buf[0] = float(np.abs(self.D).sum(axis=1).max())
        self.iter, self.done_flag, self.converged, self.sv = 0, 0, False, 10

    def init_finish(self):
        G = self.sum_buf[:self.ngram].numpy().reshape(self.n, self.n)
        self.norm_two = float(np.sqrt(np.linalg.eigvalsh(G)[-1]))
        self.normD2 = float(np.trace(G))
        dual = max(self.norm_two, float(self.max_buf[0]) / self.lam)
        self.Y = self.D / dual
        self.S = np.zeros_like(self.D)
        self.L = np.zeros_like(self.D)
        self.mu = self.mu_scale / self.norm_two

    def gram(self):
        if self.done_flag: return
        self.W = self.D - self.S + self.Y / self.mu
        self.sum_buf[:self.ngram] = torch.from_numpy((self.W.T @ self.W).ravel())

    def solve(self):
        if self.done_flag: return
        G = self.sum_buf[:self.ngram].numpy().reshape(self.n, self.n)
        w, v = np.linalg.eigh(G)
        s = np.sqrt(np.maximum(w[::-1][:self.sv], 0)); v = v[:, ::-1][:, :self.sv]
        self.iter += 1
        self.svp, self.sv = O.rank_logic(s, self.sv, self.mu, self.d)
        if self.labels is not None and self.svp == 0:              # group_sparse_RPCA.py:91-93: stop before L / S are touched
            self.done_flag = 3
            return
        self.P = (v[:, :self.svp] * (1 - 1 / (self.mu * s[:self.svp]))) @ v[:, :self.svp].T

    def shrink(self):
        if self.done_flag: return
        self.L = self.W @ self.P
        G_S = self.D - self.L + self.Y / self.mu
        self.S = O.prox_flat(G_S, self.lam / self.mu, self.groups)
        Z = self.D - self.L - self.S
        self.Y = self.Y + self.mu * Z
        self.sum_buf[self.ngram:self.ngram + 4] = torch.tensor([float((Z * Z).sum()), float(np.count_nonzero(self.S)), 0.0, 0.0],
                                                              dtype=torch.float64)
        self.mu *= self.rho                      # like the device solver: mu advances with the shrink pass (local data only)

    def shrink_a(self):
        if self.done_flag: return
        self.L = self.W @ self.P
        self.G_S = np.asfortranarray(self.D - self.L + self.Y / self.mu)
        if self.graph_cols is not None:
            self.S = np.asfortranarray(self.S)
            return
        sums = np.zeros((self.n, self.nlab))
        for f in range(self.n):
            np.add.at(sums[f], self.labels[f], self.G_S[:, f] ** 2)
        self.bsums[:] = torch.from_numpy(sums.ravel())

    def shrink_b(self):
        if self.done_flag: return
        if self.graph_cols is not None:          # S has been filled by the driver (prox on whole frames)
            Z = self.D - self.L - self.S
            self.Y = self.Y + self.mu * Z
            self.sum_buf[self.ngram:self.ngram + 4] = torch.tensor([float((Z * Z).sum()), float(np.count_nonzero(self.S)), 0.0, 0.0],
                                                                  dtype=torch.float64)
            self.mu *= self.rho
            return
        sums = self.bsums.numpy().reshape(self.n, self.nlab)
        S = np.zeros_like(self.G_S)
        with np.errstate(divide='ignore', invalid='ignore'):
            for f in range(self.n):
                eps = np.array([self.nbl] + list(self.lams[f]) + [0.0] * (self.nlab - 1 - len(self.lams[f]))) / self.mu
                nrm = np.sqrt(sums[f])
                fac = np.where(nrm > 0, np.maximum(1 - eps / nrm, 0), 0.0)
                S[:, f] = fac[self.labels[f]] * self.G_S[:, f]
        self.S = S
        Z = self.D - self.L - self.S
        self.Y = self.Y + self.mu * Z
        self.sum_buf[self.ngram:self.ngram + 4] = torch.tensor([float((Z * Z).sum()), float(np.count_nonzero(self.S)), 0.0, 0.0],
                                                              dtype=torch.float64)
        self.mu *= self.rho

    def finish_iter(self):
        """Stop test of the iteration whose scalars arrived with the last all-reduce (the Gram of the NEXT iteration has
        already been formed by then, exactly as in the device solver)."""
        if self.done_flag: return
        err = float(np.sqrt(float(self.sum_buf[self.ngram]) / self.normD2))
        self.log.append((self.iter, self.svp, err))
        if err < self.tol: self.done_flag, self.converged = 1, True
        elif self.iter >= self.max_iter: self.done_flag = 2

    def done(self): return self.done_flag != 0
    def done_by(self, k): return self.done_flag != 0 and self.iter <= k
    def finalize(self): pass

    def mask_stats(self, phase):
        if phase == 0:
            self.max_buf.zero_(); self.max_buf[1] = float(np.abs(self.S).max())
        else:
            A = np.abs(self.S)
            dl = np.abs(self.D - self.L) * (A < 0.5 * float(self.max_buf[1]))
            pos = dl[dl > 0]
            self.sum_buf[self.ngram + 4:self.ngram + 8] = torch.tensor([pos.size, pos.sum(), (pos ** 2).sum(), 0.0], dtype=torch.float64)

    def mask_device(self, sigmas=2.0):
        cnt, s1, s2 = [float(x) for x in self.sum_buf[self.ngram + 4:self.ngram + 7]]
        mean = s1 / cnt
        th = mean + sigmas * np.sqrt(max(s2 / cnt - mean * mean, 0.0))
        return np.abs(self.S) > th
