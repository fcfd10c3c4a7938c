"""Per-epoch validation metrics: the reference's `Evaluator` (MuRaL/evaluation/evaluation.py:497-588) on device tables.

The reference runs pandas group-bys and a per-row python loop over the whole validation set every epoch
(MuRaL/training.py:488-520).  Here every metric is ONE pass of a CUDA reduction over the sites (csrc/metrics.cu,
`mural_kmer_group_stats`, `mural_window_runs`) into a small integer table; only the Pearson correlations over those tables
(a few hundred numbers) are formed on the host, in float64.

Same names, arguments and printed lines as the reference: `freq_kmer_comp_multi`, `corr_calc_sub`, `calc_avg_prob`,
`Evaluator(data_local, y_prob, n_class, calibra, printer)` with `evaluate_kmer`, `evaluate_regional_score`,
`evaluate_regional_corr`.  Inputs may be the reference's pandas frames (uploaded once) or device-resident `EvalData`.

Numerics: counts are exact; probability sums are exact to 2^-37 per site (fixed point) whatever the order of the atomics.
The reference, through pandas, averages float32 probability columns IN float32 (Kahan / sequential); `EvalData.f32` rounds
the exact means to float32 to stay within one float32 ulp of those.  Float64 (calibrated) probabilities agree to ~1e-9.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

SCALE = 68719476736.0   # MURAL_METRIC_SCALE


class EvalData:
    """Device-resident validation set: order-1 local codes [n, n_cols] int64, meta (label in MURAL_META) int32,
    probabilities [n, n_class] float64; optionally chrom index / start for the regional correlations."""

    def __init__(self, flank, meta, prob, f32=False, chrom_names=None, start=None):
        self.flank = flank.contiguous()
        self.meta = meta.contiguous()
        self.f32 = bool(f32 or prob.dtype == torch.float32)
        self.prob = prob.to(torch.float64).contiguous()
        self.n, self.n_class = self.prob.shape
        self.chrom_names = chrom_names
        self.start = start
        assert self.flank.dtype == torch.int64 and self.meta.dtype == torch.int32 and self.flank.shape[0] == self.n == self.meta.shape[0]

    @staticmethod
    def from_frame(data_and_prob, n_class, device="cuda"):
        """From the reference's `data_and_prob` frame: us*/mid/ds* columns, `mut_type`, `prob0..` (evaluation.py:508-511)."""
        cols = [c for c in data_and_prob.columns if c.startswith("us")]
        R = len(cols)
        us, ds = ["us%d" % i for i in range(R, 0, -1)], ["ds%d" % i for i in range(1, R + 1)]
        if "mid" in data_and_prob.columns:                     # snv header (get_local_header, preprocessing.py:358-375)
            flank_np = data_and_prob[us + ["mid"] + ds].to_numpy(dtype=np.int64)
        else:                                                  # indel header has no centre column: a dummy one keeps us_j / ds_j at mid -/+ j
            flank_np = np.concatenate([data_and_prob[us].to_numpy(dtype=np.int64), np.zeros((len(data_and_prob), 1), np.int64),
                                       data_and_prob[ds].to_numpy(dtype=np.int64)], 1)
        flank = torch.from_numpy(np.ascontiguousarray(flank_np)).to(device)
        label = data_and_prob["mut_type"].to_numpy().astype(np.int64)
        prob = data_and_prob[["prob%d" % i for i in range(n_class)]].to_numpy()
        meta = torch.from_numpy(((label & 0x7f) << 1).astype(np.int32)).to(device)
        return EvalData(flank, meta, torch.from_numpy(np.ascontiguousarray(prob)).to(device), f32=prob.dtype == np.float32)

    def labels_host(self):
        return ((self.meta.cpu().numpy() >> 1) & 0x7f).astype(np.int64)


def _pearson(x, y):
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    if len(x) < 2:
        return float("nan")
    xm, ym = x - x.mean(), y - y.mean()
    den = np.sqrt((xm * xm).sum() * (ym * ym).sum())
    return float((xm * ym).sum() / den) if den > 0 else float("nan")


def _means(sums, cnt, f32):
    m = sums.astype(np.float64) / SCALE / cnt
    return m.astype(np.float32).astype(np.float64) if f32 else m


def kmer_group_table(ed, k, region_size=0):
    """int64 [n_regions, 5^(2*(k//2)), 1 + 2*n_class] on the host (sites, sites per label, fixed-point prob sums)."""
    d = k // 2
    if not (1 <= d <= 4 and 2 * d + 1 <= ed.flank.shape[1]):
        raise ValueError("ValueError: k-mer length %d does not fit the %d local columns" % (k, ed.flank.shape[1]))
    n_regions = 1 if region_size == 0 else ed.n // region_size
    G, W = 5 ** (2 * d), 1 + 2 * ed.n_class
    table = torch.empty((max(n_regions, 0), G, W), dtype=torch.int64, device=ed.prob.device)
    if ed.n == 0 or n_regions == 0:
        return table.zero_().cpu().numpy()
    _lib.check(_lib.lib().mural_kmer_group_stats(_lib.ptr(ed.flank), ed.n, ed.flank.shape[1], k, _lib.ptr(ed.meta), _lib.ptr(ed.prob),
                                                 ed.n_class, region_size, _lib.ptr(table), _lib.current_stream()))
    return table.cpu().numpy()


def _kmer_corr_regions(tabs, n_class, f32):
    """Correlation of observed and predicted group means over the OBSERVED groups (evaluation.py:58-65) of every region of
    `tabs` [R, G, W] at once: float64 [R, n_class]; NaN where a region has fewer than two groups or a constant column."""
    cnt = tabs[:, :, 0].astype(np.float64)
    seen = cnt > 0
    safe = np.where(seen, cnt, 1.0)
    m = seen.sum(1).astype(np.float64)
    out = np.full((tabs.shape[0], n_class), np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n_class):
            obs = tabs[:, :, 1 + i] / safe
            pred = _means(tabs[:, :, 1 + n_class + i], safe, f32)
            xm = np.where(seen, obs - (obs * seen).sum(1, keepdims=True) / m[:, None], 0.0)
            ym = np.where(seen, pred - (pred * seen).sum(1, keepdims=True) / m[:, None], 0.0)
            den = np.sqrt((xm * xm).sum(1) * (ym * ym).sum(1))
            out[:, i] = np.where((m >= 2) & (den > 0), (xm * ym).sum(1) / den, np.nan)
    return out


def _kmer_corr(tab, n_class, f32):
    return [float(c) for c in _kmer_corr_regions(tab[None], n_class, f32)[0]]


def _as_eval_data(data, n_class):
    return data if isinstance(data, EvalData) else EvalData.from_frame(data, n_class)


def freq_kmer_comp_multi(data_and_prob, k, n_class):
    """evaluation.py:48-67: per mutation subtype, the correlation over k-mer contexts of the observed rate and the mean
    predicted probability."""
    ed = _as_eval_data(data_and_prob, n_class)
    return _kmer_corr(kmer_group_table(ed, k)[0], n_class, ed.f32)


def calc_avg_prob(df, n_class):
    """evaluation.py:196-203: observed class frequencies followed by mean predicted probabilities."""
    ed = _as_eval_data(df, n_class)
    tab = kmer_group_table(ed, 3)[0].sum(0)
    cnt = float(tab[0])
    return [float(tab[1 + i] / cnt) for i in range(n_class)] + [float(_means(tab[1 + n_class + i], cnt, ed.f32)) for i in range(n_class)]


def window_table(ed, window, order=None):
    """avg_obs / avg_pred per run of consecutive sites (in `order`) sharing (chrom, start // window): [runs, 2*n_class]."""
    n_runs = C.c_int64(0)
    L = _lib.lib()
    args = (_lib.ptr(ed.meta), _lib.ptr(ed.start), _lib.ptr(order), _lib.ptr(ed.prob), ed.n, ed.n_class, int(window))
    _lib.check(L.mural_window_runs(*args, C.byref(n_runs), None, 0, _lib.current_stream()))
    rows = torch.empty((n_runs.value, 1 + 2 * ed.n_class), dtype=torch.int64, device=ed.prob.device)
    if n_runs.value:
        _lib.check(L.mural_window_runs(*args, C.byref(n_runs), _lib.ptr(rows), rows.shape[0], _lib.current_stream()))
    rows = rows.cpu().numpy()
    cnt = rows[:, 0].astype(np.float64)
    K = ed.n_class
    tab = np.empty((len(cnt), 2 * K))
    for i in range(K):
        tab[:, i] = rows[:, 1 + i] / cnt
        tab[:, K + i] = _means(rows[:, 1 + K + i], cnt, ed.f32)
    return tab


def _std(x):   # pandas Series.std: ddof = 1
    return float(np.std(x, ddof=1)) if len(x) > 1 else float("nan")


def corr_calc_sub(data, window, prob_names, order=None):
    """evaluation.py:124-193: correlation of observed and predicted window averages, windows = runs of consecutive rows with
    the same (chrom, start // window).  `data`: the reference's frame (chrom, start, mut_type, prob*; already sorted) or an
    `EvalData` with `.start` (and `order`, a device int64 permutation, when its rows are not in sorted order)."""
    n_class = len(prob_names)
    if not isinstance(data, EvalData):
        names, chrom = np.unique(data["chrom"].to_numpy().astype(str), return_inverse=True)
        label = data["mut_type"].to_numpy().astype(np.int64)
        prob = data[list(prob_names)].to_numpy()
        meta = torch.from_numpy((((label & 0x7f) << 1) | (chrom.astype(np.int64) << 8)).astype(np.int32)).cuda()
        dummy = torch.zeros((len(label), 1), dtype=torch.int64, device="cuda")
        data = EvalData(dummy, meta, torch.from_numpy(np.ascontiguousarray(prob)).cuda(), f32=prob.dtype == np.float32,
                        start=torch.from_numpy(data["start"].to_numpy().astype(np.int32)).cuda())
    tab = window_table(data, window, order)
    corr_list = []
    for i in range(n_class):
        obs, pred = tab[:, i], tab[:, n_class + i]
        if np.sum((obs == 0) | (obs == 1)) / tab.shape[0] > 0.5:
            print('Warning: too many zeros/ones (>50%) in the obs windows of size', window, 'subtype', i)
        print('CV for ', str(window) + 'bp:', _std(obs) / obs.mean(), _std(pred) / pred.mean())
        if tab.shape[0] >= 3:
            corr = _pearson(obs, pred)
        else:
            corr = 0
            print('Warning: too few windows for calculating correlation', window, 'subtype', i)
        corr_list.append(corr)
    return corr_list


class Evaluator:
    """evaluation.py:497-588.  `data_local`: the reference's frame (us*/mid/ds* + mut_type) with `y_prob` [n, n_class]
    (numpy, float32 softmax output or float64 calibrated), or an `EvalData` (then `y_prob` may be None)."""

    def __init__(self, data_local, y_prob, n_class, calibra='no_calibra', printer=print):
        self.n_class = n_class
        self.prob_names = ['prob' + str(i) for i in range(n_class)]
        self.data_local = data_local
        self.y_prob = y_prob
        self.printer = printer
        self.calibra = calibra
        self.data_and_prob = self.preprocess()
        self.kmer_out_identify, self.regional_out_identify = self.set_output_identifiers()
        self.metrics = {}

    def preprocess(self):
        if isinstance(self.data_local, EvalData):
            self.ed = self.data_local
            return None
        import pandas as pd
        y_prob = pd.DataFrame(data=np.copy(self.y_prob), columns=self.prob_names)
        data_and_prob = pd.concat([self.data_local.reset_index(drop=True), y_prob], axis=1)
        self.ed = EvalData.from_frame(data_and_prob, self.n_class)
        return data_and_prob

    def set_output_identifiers(self):
        kmer = {'no_calibra': 'mer correlation - all: ', 'FullDiri': 'mer correlation(after fdiri_cal)',
                'Poisson': 'mer correlation(after Poisson_cal)'}
        regional = {'no_calibra': 'regional corr (validation):', 'FullDiri': 'regional corr (validation, after fdiri_cal):',
                    'Poisson': 'regional corr (validation, after Poisson_cal):'}
        return kmer[self.calibra], regional[self.calibra]

    def evaluate_kmer(self, kmer_list=[3, 5, 7]):
        if self.calibra == 'no_calibra' and self.data_and_prob is not None:
            self.printer("valid_data_and_prob.iloc[0:10]", self.data_and_prob.iloc[0:10])
        for k in kmer_list:
            kmer_corr = freq_kmer_comp_multi(self.ed, k, self.n_class)
            self.printer(f"{k}{self.kmer_out_identify}", kmer_corr)
            self.metrics['kmer%d' % k] = kmer_corr

    def evaluate_regional_corr(self, chr_pos, win_size_list=[100000, 500000], save_valid_preds=False, save_path=None):
        """chr_pos: frame with chrom, start, end, strand in the row order of the predictions (get_position_info).  Rows are
        ordered by chrom NAME then start as `sort_values(['chrom', 'start'])` does (evaluation.py:526)."""
        names, chrom = np.unique(chr_pos.iloc[:, 0].to_numpy().astype(str), return_inverse=True)
        start = chr_pos.iloc[:, 1].to_numpy().astype(np.int64)
        order = np.lexsort((start, chrom))
        if not np.array_equal(order, np.arange(len(order))):
            # sort_values' default quicksort is not stable for ties; ties (same chrom and start) fall in the same window anyway
            order_dev = torch.from_numpy(order).to(self.ed.prob.device)
        else:
            order_dev = None
        label = self.ed.labels_host()
        ed = EvalData(self.ed.flank, torch.from_numpy((((label & 0x7f) << 1) | (chrom.astype(np.int64) << 8)).astype(np.int32)).to(self.ed.prob.device),
                      self.ed.prob, f32=self.ed.f32, start=torch.from_numpy(start.astype(np.int32)).to(self.ed.prob.device))
        for win_size in win_size_list:
            corr_win = corr_calc_sub(ed, win_size, self.prob_names, order_dev)
            self.printer(self.regional_out_identify, str(win_size) + 'bp', corr_win)
            self.metrics['window%d' % win_size] = corr_win
        if save_valid_preds:
            import pandas as pd
            if self.data_and_prob is not None:
                cols = self.data_and_prob[['mut_type'] + self.prob_names]
            else:       # device-resident input: bring labels and probabilities back once for the file
                cols = pd.DataFrame(self.ed.prob.cpu().numpy(), columns=self.prob_names)
                cols.insert(0, 'mut_type', self.ed.labels_host())
            df = pd.concat((chr_pos.reset_index(drop=True), cols), axis=1)
            df.columns = ['chrom', 'start', 'end', 'strand', 'mut_type'] + self.prob_names
            df.sort_values(['chrom', 'start'], inplace=True)
            df.to_csv(save_path + '.valid_preds.tsv.gz', sep='\t', float_format='%.4g', index=False)

    def evaluate_regional_score(self, valid_size, kmer_list=[3, 5]):
        """evaluation.py:545-588: sum over regions of consecutive sites of (1 - corr)^2 of the first two k-mer lengths, and the
        correlation of the regions' observed / predicted class averages."""
        if valid_size > 10000 * 10:
            region_size = 10000
        else:
            region_size = valid_size // 10
        n_regions = valid_size // region_size
        self.printer('n_regions:', n_regions)
        # the reference slices iloc[region_size*i : region_size*(i+1)] of the frame: regions cover the first valid_size rows
        sub = self.ed if valid_size >= self.ed.n else EvalData(self.ed.flank[:valid_size], self.ed.meta[:valid_size],
                                                               self.ed.prob[:valid_size], f32=self.ed.f32)
        tabs = [kmer_group_table(sub, k, region_size)[:n_regions] for k in kmer_list[:2]]
        K = self.n_class
        score = 0
        for t in tabs:      # the reference adds (1 - corr)^2 of every class, k-mer length and region; NaN propagates as it does there
            score += float(np.sum((1 - _kmer_corr_regions(t, K, self.ed.f32)) ** 2))
        tot = tabs[0].sum(1).astype(np.float64)                 # [n_regions, W]: calc_avg_prob of every region
        cnt = tot[:, :1]
        region_avg = np.concatenate([tot[:, 1:1 + K] / cnt, _means(tabs[0].sum(1)[:, 1 + K:], cnt, self.ed.f32)], 1).reshape(n_regions, 2 * K)
        corr_list = [_pearson(region_avg[:, i], region_avg[:, i + K]) for i in range(K)]
        corr_list_perfix = {'no_calibra': 'corr_list: ', 'FullDiri': 'corr_list(after fdiri_cal)', 'Poisson': 'corr_list(after Poisson_cal)'}
        regional_score_perfix = {'no_calibra': 'regional score: ', 'FullDiri': 'regional score(after fdiri_cal)',
                                 'Poisson': 'regional score(after Poisson_cal)'}
        self.printer(corr_list_perfix[self.calibra], corr_list)
        self.printer(regional_score_perfix[self.calibra], score, n_regions)
        self.metrics['score'] = score
        self.metrics['corr_list'] = corr_list
