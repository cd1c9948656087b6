"""Ensemble diagnostics of the reference, restated (diagnostics/MeanAndVariance.h, ChainStats.h, MultiChainStats.h;
formulas: doc/ABMCMC.tex:910-952, i.e. Gelman et al. BDA3 ch. 11), plus a reader for the reference's ``.stat`` result
archives (Boost binary archive v15; field order MultiChainStats.h:221-224, ChainStats.h:85-87, MeanAndVariance.h:78-80,
MCMCStatistics.h:32-34) so shipped results can be re-analysed without Boost.

Inputs are per-sequence sufficient statistics, which is what the GPU sampler returns (``abm_chains_get_statistics``,
``abm_chains_get_series``) and what an NCCL all-reduce can combine across GPUs.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class ChainStats:
    """One sample sequence: MeanAndVariance (n, sum, sumOfSquares), variogram, mean end state, MCMC counters."""
    n_samples: int
    sum: np.ndarray
    sum_sq: np.ndarray
    vario: np.ndarray = None          # [nLags][d]
    vario_stride: int = 1
    mean_end_state: np.ndarray = None
    n_accepted: np.ndarray = None     # [2][2]
    n_proposals: np.ndarray = None

    def mean(self):  # MeanAndVariance.h:52-54
        return self.sum * (1.0 / self.n_samples)

    def sample_variance(self):  # MeanAndVariance.h:57-59
        return (self.sum_sq - self.sum * self.sum * (1.0 / self.n_samples)) * (1.0 / (self.n_samples - 1))

    @staticmethod
    def from_series(series: np.ndarray, n_lags: int = 200, max_lag_proportion: float = 0.5, **kw) -> "ChainStats":
        """ChainStats(synopsisSamples, nLags, maxLagProportion, ...) (ChainStats.h:27-40)."""
        x = np.asarray(series, dtype=np.float64)
        return ChainStats(n_samples=x.shape[0], sum=x.sum(axis=0), sum_sq=(x * x).sum(axis=0),
                          vario=variogram(x, n_lags, max_lag_proportion),
                          vario_stride=int((max_lag_proportion * x.shape[0]) / (n_lags - 1.0)), **kw)


def variogram(samples: np.ndarray, n_lags: int = 100, max_lag_proportion: float = 0.5) -> np.ndarray:
    """ChainStats::variogram (ChainStats.h:66-79): V_t = mean_i (x_i - x_{i+t})^2 at lags t = j * stride."""
    x = np.asarray(samples, dtype=np.float64)
    n = x.shape[0]
    stride = int((max_lag_proportion * n) / (n_lags - 1.0))
    out = np.zeros((n_lags, x.shape[1]))
    for j in range(n_lags):
        t = j * stride
        d = x[: n - t] - x[t:]
        out[j] = (d * d).sum(axis=0) / (n - t)
    return out


@dataclass
class MultiChainStats:
    """MultiChainStats (diagnostics/MultiChainStats.h:15-227) over a list of equally long sequences."""
    chains: List[ChainStats] = field(default_factory=list)
    kappa: float = 0.0
    description: str = ""
    exec_time_ms: int = 0
    cpuinfo: str = ""

    def n_chains(self):
        return len(self.chains)

    def n_samples_per_chain(self):
        return self.chains[0].n_samples if self.chains else 0

    def mean(self):  # :183-190
        return sum(c.mean() for c in self.chains) / self.n_chains()

    def W(self):  # :174-181
        return sum(c.sample_variance() for c in self.chains) / self.n_chains()

    def B(self, reference_integer_division: bool = True):
        """:160-169.  The reference multiplies by ``nSamplesPerChain()/(nChains()-1)`` in INTEGER arithmetic (:167);
        keep that to reproduce its published numbers, or pass False for the textbook n/(m-1)."""
        mu = self.mean()
        var = sum((c.mean() - mu) ** 2 for c in self.chains)
        n, m = self.n_samples_per_chain(), self.n_chains()
        return var * (n // (m - 1) if reference_integer_division else n / (m - 1))

    def var_plus(self, **kw):  # :151-153
        n = self.n_samples_per_chain()
        return self.W() * ((n - 1.0) / n) + self.B(**kw) * (1.0 / n)

    def potential_scale_reduction(self, **kw):  # :141-143  (R-hat)
        n = self.n_samples_per_chain()
        return np.sqrt(self.B(**kw) / (self.W() * (1.0 * n)) + (n - 1.0) / n)

    def mean_variogram(self):  # :124-133
        return sum(c.vario for c in self.chains) / self.n_chains()

    def autocorrelation(self, **kw):  # :113-121
        return 1.0 - self.mean_variogram() / (2.0 * self.var_plus(**kw))

    def effective_samples(self, **kw):
        """:81-102: n_eff = m n / (1 + 2 stride sum_t rho_t), rho_t = 1 - V_t / (2 var+), the sum running while the
        running minimum of rho stays positive."""
        V = self.mean_variogram()
        vp = self.var_plus(**kw)
        d = V.shape[1]
        out = np.zeros(d)
        for k in range(d):
            t = 0
            next_s = 1.0 - V[0][k] / (2.0 * vp[k])
            sumt = 0.0
            while next_s > 0.0:
                sumt += next_s
                t += 1
                next_s = min(next_s, (1.0 - V[t][k] / (2.0 * vp[k])) if t < V.shape[0] else -1.0)
            out[k] = self.n_samples_per_chain() * self.n_chains() / (1.0 + 2.0 * self.chains[0].vario_stride * sumt)
        return out

    def mean_end_state(self):  # :71-78
        return sum(c.mean_end_state for c in self.chains) / self.n_chains()


def rhat_from_moments(n: np.ndarray, s: np.ndarray, sq: np.ndarray):
    """R-hat, W, B, var+ from per-sequence (n, sum, sum of squares) with the textbook n/(m-1) factor; sequences may
    have different lengths only in the sense that the mean length is used (the GPU ensemble uses equal lengths).
    n: [m], s, sq: [m][d]."""
    n = np.asarray(n, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    sq = np.asarray(sq, dtype=np.float64)
    m = n.size
    nbar = n.mean()
    means = s / n[:, None]
    var = (sq - s * s / n[:, None]) / (n[:, None] - 1.0)
    W = var.mean(axis=0)
    B = nbar / (m - 1.0) * ((means - means.mean(axis=0)) ** 2).sum(axis=0)
    var_plus = W * (nbar - 1.0) / nbar + B / nbar
    return np.sqrt(var_plus / W), W, B, var_plus


def effective_samples_from_variogram(V: np.ndarray, var_plus: np.ndarray, n_per_chain: float, n_chains: float, stride: int):
    """MultiChainStats::effectiveSamples (:81-102) from a mean variogram V [lags][d] at lags 0, stride, 2 stride, ...:
    n_eff = m n / (1 + 2 stride sum_t rho_t), rho_t = 1 - V_t / (2 var+), summed while the running minimum of rho is positive.
    Returns (n_eff [d], rho [lags][d])."""
    V = np.asarray(V, dtype=np.float64)
    rho = 1.0 - V / (2.0 * np.asarray(var_plus, dtype=np.float64))
    n_lags, d = V.shape
    n_eff = np.zeros(d)
    for q in range(d):
        t, next_s, sumt = 0, 1.0 - V[0][q] / (2.0 * var_plus[q]), 0.0
        while next_s > 0.0:
            sumt += next_s
            t += 1
            next_s = min(next_s, rho[t][q] if t < n_lags else -1.0)
        n_eff[q] = n_per_chain * n_chains / (1.0 + 2.0 * stride * sumt)
    return n_eff, rho


def ensemble_effective_samples(series: np.ndarray, thin: int, n: np.ndarray, s: np.ndarray, sq: np.ndarray,
                               max_lag_proportion: float = 0.5, group: int = 1):
    """Effective sample size of MANY equally long chains, with the reference's estimator
    (MultiChainStats::effectiveSamples, diagnostics/MultiChainStats.h:81-102; BDA3 ch. 11):

        n_eff = m n / (1 + 2 stride sum_t rho_t),   rho_t = 1 - V_t / (2 var+),   t = 0, stride, 2 stride, ...

    summed while the running minimum of rho_t stays positive.  The reference evaluates the variogram V_t at 200 strided
    lags from the full series of each sequence (ChainStats.h:66-79); with thousands of chains the full series cannot be
    kept, but they are not needed either: V_t = E (x_i - x_{i+t})^2 is estimated from every `thin`-th sample of every chain
    (`series`, [m][k][d], stride = thin), each lag averaged over all chains and all pairs, which with m in the thousands is
    far more precise than the reference's per-sequence average.  var+ = (n-1)/n W + B/n from the exact per-chain moments
    (n, s, sq as `rhat_from_moments`).  The estimate is valid when the window (n samples per chain) is long against the
    autocorrelation time; the lags run to max_lag_proportion * n as in the reference (FiguresForPaper.h:89).

    group > 1 additionally returns the R-hat of `m / group` super-sequences, each the pooled samples of `group` chains:
    for stationary chains R-hat^2 = 1 + 1 / (effective samples per sequence), so sequences of a few autocorrelation times
    cannot fall below ~1.05 however well the ensemble has converged; pooled sequences can.

    Returns a dict: n_eff [d] (all chains), tau [d] (samples per effective sample), rho [lags][d], rhat [d], rhat_grouped [d],
    W, B, var_plus."""
    x = np.asarray(series, dtype=np.float64)
    m, k, d = x.shape
    n = np.asarray(n, dtype=np.float64)
    rhat, W, B, var_plus = rhat_from_moments(n, s, sq)
    nbar = float(n.mean())
    n_lags = max(2, min(k, int(max_lag_proportion * nbar / thin) + 1))
    V = np.zeros((n_lags, d))
    for j in range(1, n_lags):
        diff = x[:, : k - j] - x[:, j:]
        V[j] = (diff * diff).mean(axis=(0, 1))
    n_eff, rho = effective_samples_from_variogram(V, var_plus, nbar, m, thin)
    out = {"n_eff": n_eff, "tau": nbar * m / n_eff, "rho": rho, "rhat": rhat, "W": W, "B": B, "var_plus": var_plus,
           "lags": n_lags, "stride": thin}
    if group > 1:
        mg = (m // group) * group
        ng = n[:mg].reshape(-1, group).sum(axis=1)
        sg = np.asarray(s, dtype=np.float64)[:mg].reshape(-1, group, d).sum(axis=1)
        qg = np.asarray(sq, dtype=np.float64)[:mg].reshape(-1, group, d).sum(axis=1)
        out["rhat_grouped"] = rhat_from_moments(ng, sg, qg)[0]
    return out


# ------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, raw: bytes):
        self.raw, self.pos = raw, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.raw, self.pos)
        self.pos += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def skip(self, n):
        self.pos += n

    def string(self):
        n = self.take("Q")
        s = self.raw[self.pos:self.pos + n]
        self.pos += n
        return s.decode("latin1")

    def f64s(self):
        n = self.take("Q")
        a = np.frombuffer(self.raw, dtype="<f8", count=n, offset=self.pos).copy()
        self.pos += 8 * n
        return a


def read_stat(path) -> MultiChainStats:
    """Reads a reference ``.stat`` archive (boost::archive::binary_oarchive, archive version 15, little endian).

    Layout (decoded from the shipped files): signature string, u16 version, 4 size bytes, class
    header of MultiChainStats, cpuinfo, kappa, description, execTimeMilliSeconds, vector<ChainStats> header, then per
    sequence nSamples, sum, sumOfSquares, nLags x variogram row, varioStride, meanEndState, nAccepted, nProposals.  Class
    headers appear once, in front of the first object of each class.
    """
    r = _Reader(open(path, "rb").read())
    if r.string() != "serialization::archive":
        raise ValueError("not a boost binary archive")
    version = r.take("H")
    r.skip(4)                      # sizeof(int), sizeof(long), sizeof(float), sizeof(double)
    r.skip(4 + 5)                  # tracking/class header of the top-level object
    out = MultiChainStats()
    out.cpuinfo = r.string()
    out.kappa = r.take("d")
    out.description = r.string()
    out.exec_time_ms = r.take("q")
    r.skip(5)                      # class header of std::vector<ChainStats>
    n_seq = r.take("Q")
    r.skip(4)                      # item_version
    first = True
    for _ in range(n_seq):
        if first:
            r.skip(5)              # ChainStats class header
            r.skip(5)              # MeanAndVariance class header
        n = r.take("i")
        s = r.f64s()
        sq = r.f64s()
        if first:
            r.skip(5)              # valarray<valarray<double>> class header
        n_lags = r.take("Q")
        vario = np.stack([r.f64s() for _ in range(n_lags)])
        stride = r.take("i")
        end = r.f64s()
        if first:
            r.skip(5)              # MCMCStatistics class header
        acc = np.array([[0, 0], [0, 0]])
        prop = np.array([[0, 0], [0, 0]])
        for arr in (acc, prop):
            if r.take("Q") != 2:
                raise ValueError("unexpected counter layout")
            for a in range(2):
                if r.take("Q") != 2:
                    raise ValueError("unexpected counter layout")
                arr[a, 0], arr[a, 1] = r.take("i"), r.take("i")
        out.chains.append(ChainStats(n, s, sq, vario, stride, end, acc, prop))
        first = False
    out.version = version
    return out
