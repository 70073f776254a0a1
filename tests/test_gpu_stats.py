"""GPU tests of the reductions: device_process.pmean/pvar/pstd and montecarlo
(moments within 1e-12 relative of the reference's values, histogram counts
and out-of-range counts bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import sde_oracle as orc
from tests.cases import golden

pytestmark = pytest.mark.gpu


def sd():
    import sdepy_b200
    return sdepy_b200


def test_psummaries_match_reference():
    m = sd()
    g = golden('stats_lognorm')
    dp = m.device_process(g['t'], torch.from_numpy(g['x']).cuda())
    assert np.allclose(np.asarray(dp.pmean()), g['pmean'], rtol=1e-14, atol=0)
    assert np.allclose(np.asarray(dp.pvar()), g['pvar'], rtol=1e-12, atol=0)
    assert np.allclose(np.asarray(dp.pstd()), g['pstd'], rtol=1e-12, atol=0)
    assert np.allclose(np.asarray(dp.pvar(ddof=1)), g['pvar1'], rtol=1e-12, atol=0)
    assert dp.pmean().shape == g['pmean'].shape
    assert np.array_equal(np.asarray(dp.pmin())[..., 0], g['x'].min(axis=-1))
    assert np.array_equal(np.asarray(dp.pmax())[..., 0], g['x'].max(axis=-1))
    assert np.array_equal(np.asarray(dp.cpu()), g['x'])


@pytest.mark.parametrize('where', ['host', 'device'])
def test_montecarlo_one_shot_and_chunked(where):
    m = sd()
    g = golden('stats_lognorm')
    xT = g['x'][-1]

    def put(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda() if where == 'device' else a
    one = m.montecarlo(put(xT), bins=25)
    chunk = m.montecarlo(bins=25)
    for c in (xT[:, :1500], xT[:, 1500:2700], xT[:, 2700:]):
        chunk.update(put(c))
    for tag, a in (('one', one), ('chunk', chunk)):
        assert a.paths == xT.shape[-1]
        assert np.allclose(a.mean(), g[tag + '_mean'], rtol=1e-13, atol=0)
        assert np.allclose(a.var(), g[tag + '_var'], rtol=1e-11, atol=0)
        assert np.allclose(a.std(), g[tag + '_std'], rtol=1e-11, atol=0)
        assert np.allclose(a.stderr(), g[tag + '_stderr'], rtol=1e-11, atol=0)
        assert np.allclose(a.skew(), g[tag + '_skew'], rtol=1e-9, atol=0)
        assert np.allclose(a.kurtosis(), g[tag + '_kurt'], rtol=1e-9, atol=0)
        for i in range(2):
            counts, edges = a[i].histogram()
            assert np.array_equal(edges, g[tag + '_edges'][i])
            assert np.array_equal(counts, g[tag + '_counts'][i])
            assert a[i].outpaths == g[tag + '_outpaths'][i]


def test_histogram_semantics_large_random():
    """Half-open bins, last bin closed, values on edges, explicit edges."""
    m = sd()
    rng = np.random.default_rng(0)
    x = rng.normal(size=3_000_017)
    x[:1000] = np.linspace(-3, 3, 1000)          # many values exactly on edges
    a = m.montecarlo(x, bins=60, range=(-3., 3.))
    c, e = np.histogram(x, bins=60, range=(-3., 3.))
    counts, edges = a.histogram()
    assert np.array_equal(edges, e) and np.array_equal(counts, c)
    assert a.outpaths == x.size - c.sum()
    edges2 = np.array([-2., -1.5, -.2, 0., .1, 1.7, 4.])
    b = m.montecarlo(x, bins=edges2)
    c2, _ = np.histogram(x, bins=edges2)
    assert np.array_equal(b.histogram()[0], c2)
    o = orc.moments_histogram(bins=60, range=(-3., 3.))
    o.update(x)
    assert np.allclose(a.mean(), o.mean(), rtol=1e-10, atol=1e-15)
    assert np.allclose(a.var(), o.var(), rtol=1e-12)


def test_device_process_cdf_chf_interp_match_reference():
    """process.cdf / chf / interpolation on the HBM-resident container
    (reference infrastructure.py:544-633, 1125-1209): cdf and interpolated
    values exactly, chf within 1e-12."""
    m = sd()
    g, s = golden('stats_cdf_chf'), golden('stats_lognorm')
    dp = m.device_process(s['t'], torch.from_numpy(s['x']).cuda())
    assert np.array_equal(dp(g['tq']).cpu().numpy(), g['interp'])
    assert np.array_equal(dp(.25, .5).cpu().numpy(), g['incr'])
    assert np.array_equal(dp.cdf(g['xq']), g['cdf_tl'])
    assert np.array_equal(dp.cdf(g['tq'], g['xq']), g['cdf_t'])
    assert dp.chf(g['uq']).shape == g['chf_tl'].shape
    assert np.allclose(dp.chf(g['uq']), g['chf_tl'], rtol=1e-12, atol=1e-15)
    assert np.allclose(dp.chf(g['tq'], g['uq']), g['chf_t'], rtol=1e-12, atol=1e-15)
    with pytest.raises(TypeError):
        dp.chf()


def test_montecarlo_pdf_cdf_match_reference():
    m = sd()
    g, s = golden('stats_cdf_chf'), golden('stats_lognorm')
    a = m.montecarlo(s['x'][-1], bins=25)
    for i in range(2):
        assert np.allclose(a[i].pdf(g['xs']), g['mc_pdf'][i], rtol=1e-12)
        assert np.allclose(a[i].cdf(g['xs']), g['mc_cdf'][i], rtol=1e-12)
        assert np.allclose(a[i].pdf(g['xs'], method='interp'), g['mc_pdf_i'][i], rtol=1e-12)
        assert np.allclose(a[i].cdf(g['xs'], method='interp'), g['mc_cdf_i'][i], rtol=1e-12)
    with pytest.raises(ValueError):
        a[0].pdf(1., method='nope')


def test_cdf_chf_of_a_simulated_process_against_closed_forms():
    """Wiener process: cdf = Phi(x/sqrt(t)), chf = exp(-u^2 t/2) within Monte
    Carlo error (the quant tests of the reference, tests/test_quant.py:305-351)."""
    import scipy.stats
    m = sd()
    paths = 400_000
    x = m.wiener_process(paths=paths, steps=41, seed=6, output='device')(np.linspace(0, 2, 5))
    xq, uq = np.linspace(-2, 2, 9), np.linspace(-1.5, 1.5, 7)
    t = np.array([.5, .75, 2.])
    c = x.cdf(t, xq)
    want = scipy.stats.norm.cdf(xq[None, :]/np.sqrt(t[:, None]))
    # .75 is interpolated between knots .5 and 1.: variance is not linear in t for
    # interpolated values, so compare only the knots exactly on the timeline
    for k in (0, 2):
        assert np.abs(c[k] - want[k]).max() < 4*.5/np.sqrt(paths)
    f = x.chf(t[[0, 2]], uq)
    wantf = np.exp(-uq[None, :]**2*t[[0, 2], None]/2)
    assert np.abs(f - wantf).max() < 5/np.sqrt(paths)


def test_device_process_time_and_value_summaries_bit_exact():
    """tmin .. tint and vmin .. vstd on the resident slab: the reference's own
    results (sequential NumPy reductions of a non-contiguous axis), bit for
    bit, and the results stay on the device."""
    from tests.test_oracle_golden import TIME_AXIS
    m = sd()
    g = golden('stats_time_axis')
    dp = m.device_process(g['t'], torch.from_numpy(g['x']).cuda())
    for k in TIME_AXIS:
        r = getattr(dp, k)()
        assert isinstance(r, m.device_process) and r.x.is_cuda, k
        assert r.shape == g[k].shape, (k, r.shape, g[k].shape)
        assert np.array_equal(r.x.cpu().numpy(), g[k]), k
    assert np.array_equal(dp.tvar(ddof=1).x.cpu().numpy(), g['tvar1'])
    assert np.array_equal(dp.vstd(ddof=1).x.cpu().numpy(), g['vstd1'])
    q = dp.tdiff(dt_exp=.5, fwd=False)
    assert np.array_equal(q.x.cpu().numpy(), g['tdiff_half_bwd'])
    assert np.array_equal(q.t, g['tdiff_half_bwd_t'])
    assert np.array_equal(dp.tmin().t, g['tmin_t'])
    # NaNs propagate through min / max like NumPy's
    x = g['x'].copy()
    x[5, 1, 2, 7] = np.nan
    dn = m.device_process(g['t'], torch.from_numpy(x).cuda())
    assert np.array_equal(dn.tmax().x.cpu().numpy(), x.max(axis=0, keepdims=True),
                          equal_nan=True)


def test_path_dependent_payoffs_on_the_device():
    """Asian and lookback payoffs of a simulated lognormal process computed
    from the resident slab (doc/quickguide.rst:663-669 does this on the host):
    same numbers as NumPy on the copied-back paths."""
    m = sd()
    tt = np.linspace(0., 1., 53)
    p = m.lognorm_process(x0=100., mu=.03, sigma=.2, paths=20000, steps=tt,
                          seed=9, output='device')(tt)
    host = p.cpu()
    asian = np.maximum(np.asarray(host).mean(axis=0) - 100., 0.)
    lookback = np.asarray(host).max(axis=0) - np.asarray(host)[-1]
    a = (p.tmean().x[0] - 100.).clamp_min(0.)
    lb = p.tmax().x[0] - p.x[-1]
    assert np.array_equal(a.cpu().numpy(), asian)
    assert np.array_equal(lb.cpu().numpy(), lookback)


@pytest.mark.parametrize('lo,width,nbins', [(0., 1., 7), (-3., 6., 100), (1e6, 1e-3, 64),
                                            (-1e-300, 2e-300, 10), (123.456, 7.89e5, 4000)])
def test_fused_histogram_next_to_the_bin_boundaries(lo, width, nbins):
    """mc_update_kernel decides the bin arithmetically unless the value lies
    within a rigorous margin of a boundary (then it compares with the exact
    edges): values ON every edge, one ulp below and above it, and random ones,
    at ranges whose edges are coarse or fine in ulps -- counts must equal
    numpy.histogram's for range=None (edges from the sample's min / max) and for
    an explicit range, one-shot and cumulated."""
    m = sd()
    rng = np.random.default_rng(int(nbins))
    hi = lo + width
    e = np.linspace(lo, hi, nbins + 1)
    near = np.concatenate((e, np.nextafter(e, -np.inf), np.nextafter(e, np.inf)))
    x = np.concatenate((near, rng.uniform(lo, hi, 200_003), [lo, hi]))
    x = x[(x >= lo) & (x <= hi)]
    rng.shuffle(x)
    a = m.montecarlo(x, bins=nbins)
    c, edges = np.histogram(x, bins=nbins)
    assert np.array_equal(a.histogram()[1], edges)
    assert np.array_equal(a.histogram()[0], c) and a.outpaths == 0
    # explicit range narrower than the data, then a second sample cumulated
    r = (lo + .25*width, lo + .75*width)
    b = m.montecarlo(x, bins=nbins, range=r)
    c1, e1 = np.histogram(x, bins=nbins, range=r)
    assert np.array_equal(b.histogram()[1], e1) and np.array_equal(b.histogram()[0], c1)
    assert b.outpaths == x.size - c1.sum()
    y = rng.uniform(lo, hi, 50_001)
    b.update(y)
    c2, _ = np.histogram(y, bins=e1)
    assert np.array_equal(b.histogram()[0], c1 + c2)
    assert b.outpaths == x.size + y.size - (c1 + c2).sum()
    assert np.allclose(b.mean(), np.concatenate((x, y)).mean(), rtol=1e-12)


def test_fused_histogram_degenerate_samples():
    m = sd()
    a = m.montecarlo(np.full(1000, 2.5), bins=10)              # min == max: widened by 1/2
    c, e = np.histogram(np.full(1000, 2.5), bins=10)
    assert np.array_equal(a.histogram()[0], c) and np.array_equal(a.histogram()[1], e)
    x = np.arange(10.)
    x[3] = np.nan
    with pytest.raises(ValueError):
        m.montecarlo(x, bins=5)                                  # numpy: range not finite
    b = m.montecarlo(x, bins=5, range=(0., 9.))                 # NaN counted outside
    assert b.histogram()[0].sum() == 9 and b.outpaths == 1
    one = m.montecarlo(np.array([1.5]), bins=3)
    assert one.histogram()[0].sum() == 1 and one.paths == 1


def test_cdf_any_order_and_number_of_thresholds():
    """path_cdf_kernel rank-sorts the thresholds per block: unsorted, repeated,
    infinite and NaN thresholds, and more than 2048 of them (two passes)."""
    import torch
    m = sd()
    rng = np.random.default_rng(8)
    x = rng.normal(size=(1, 300_001))
    x[0, :50] = np.nan
    dp = m.device_process(np.zeros(1), torch.from_numpy(x).cuda())
    q = np.array([.3, -1., .3, np.inf, -np.inf, np.nan, 2.5, 0., -1.])
    got = dp.cdf(q)
    want = np.array([(x[0] <= v).mean() for v in q])
    assert np.array_equal(got.reshape(-1), want)
    q = rng.normal(size=5000)
    got = dp.cdf(q).reshape(-1)
    srt = np.sort(x[0][~np.isnan(x[0])])
    want = np.searchsorted(srt, q, side='right')/x.shape[1]
    assert np.array_equal(got, want)
    # interpolated row, odd path count (scalar tail of the 16-byte loads)
    x2 = rng.normal(size=(2, 100_001))
    dp2 = m.device_process(np.array([0., 1.]), torch.from_numpy(x2).cuda())
    qq = np.linspace(-2, 2, 33)
    y = .25*x2[1] + .75*x2[0]
    assert np.array_equal(dp2.cdf(.25, qq).reshape(-1), np.array([(y <= v).mean() for v in qq]))
    u = np.linspace(-3., 3., 21)
    assert np.allclose(dp2.chf(.25, u).reshape(-1), np.exp(1j*u[:, None]*y).mean(axis=-1),
                       rtol=1e-12, atol=1e-15)
    u2 = np.array([.1, -2., .7, 5., .1])                          # not a grid: plain sincos
    assert np.allclose(dp2.chf(.25, u2).reshape(-1), np.exp(1j*u2[:, None]*y).mean(axis=-1),
                       rtol=1e-12, atol=1e-15)
