"""CPU tests: the oracle (oracle/sde_oracle.py) reproduces every golden
fixture generated from the unmodified reference BIT-EXACTLY."""
import numpy as np
import pytest

from oracle import sde_oracle as orc
from tests.cases import (golden, REPLAY, HW, hw_corr, MERTON, KOU)


@pytest.mark.parametrize('name', sorted(REPLAY))
def test_replay_bit_exact(name):
    g = golden(name)
    model, params, kw = REPLAY[name]
    where = np.searchsorted(g['grid'], g['tt'])
    # the recorded (t, dt) are what the oracle's grid walk must visit
    assert np.array_equal(g['t'], g['grid'][:-1])
    assert np.array_equal(g['dt'], g['grid'][1:] - g['grid'][:-1])
    dJ = g['dJ'] if 'dJ' in g else None
    dN = g['dN'] if 'dN' in g else None
    out, info = orc.euler_replay(model, params, kw['x0'], g['grid'], where,
                                 g['dW'], dJ, dN, y0=kw.get('y0'),
                                 full=kw.get('full', False))
    outs = out if isinstance(out, tuple) else (out,)
    for i, o in enumerate(outs):
        ref = g['out%d' % i]
        assert o.shape == ref.shape
        assert np.array_equal(o, ref), name
    for key in ('negative_y_count', 'jump_count'):
        if key in g:
            assert np.array_equal(info[key], g[key])
    assert int(g['computed_steps']) == g['grid'].size - 1
    assert int(g['stored_steps']) == g['tt'].size - 1


@pytest.mark.parametrize('name,addaxis', [('replay_system_stacked', False),
                                          ('replay_system_addaxis', True)])
def test_user_system_replay_bit_exact(name, addaxis):
    """Both stacking layouts of a user-defined q=2 system."""
    from tests.cases import user_system
    g = golden(name)
    where = np.searchsorted(g['grid'], g['tt'])
    x, y = orc.system_replay(
        user_system, 2, dict(mu=.05, sigma=g['p_sigma'], xi=g['p_xi']),
        (1., .3), g['grid'], where, g['dW'], addaxis)
    assert np.array_equal(x, g['out0']) and np.array_equal(y, g['out1'])


def test_user_system_with_jumps_replay_bit_exact():
    """'dt' + 'dn' + 'dw' terms, time-dependent coefficient (sorted-id sum)."""
    from tests.cases import jump_system, k_of_t
    g = golden('replay_system_jumps')
    where = np.searchsorted(g['grid'], g['tt'])
    x, y = orc.system_replay(jump_system, 2, dict(k=k_of_t), (1., .5), g['grid'], where,
                             g['dW'], True, dN=g['dN'])
    assert np.array_equal(x, g['out0']) and np.array_equal(y, g['out1'])


def test_step_grid_matches_reference_merge():
    for name in sorted(REPLAY):
        g = golden(name)
    g = golden('replay_oruh_ragged')
    tt, grid, where = orc.step_grid((0., .37, 1.),
                                    (0.1, .2, .21, .5, .93, 1.5))
    assert np.array_equal(grid, g['grid'])
    g = golden('replay_lognorm')
    tt, grid, where = orc.step_grid(np.linspace(0, 1, 11), 41)
    assert np.array_equal(grid, g['grid'])
    assert np.array_equal(grid[where], tt)


def test_seeded_lognorm():
    g = golden('seeded_lognorm')
    out, _ = orc.self_driven('lognorm', dict(mu=.05, sigma=.2), 1., g['tt'],
                             30, 101, np.random.default_rng(int(g['seed'])))
    assert np.array_equal(out, g['out0'])


def test_seeded_hw3_time_dependent_corr():
    g = golden('seeded_hw3_tdep')
    out, _ = orc.self_driven(
        'hull_white', dict(theta=HW['theta'], k=HW['k'], sigma=HW['sigma']),
        HW['x0'], g['tt'], 26, 101, np.random.default_rng(int(g['seed'])),
        wshape=(3,), corr=hw_corr)
    assert np.array_equal(out, g['out0'])


def test_seeded_heston():
    g = golden('seeded_heston')
    out, info = orc.self_driven(
        'heston', dict(mu=.03, sigma=1., theta=.04, k=2., xi=.3), 100.,
        g['tt'], 50, 101, np.random.default_rng(int(g['seed'])),
        wshape=(2,), corr=orc.rho_to_corr(-.7), y0=.04)
    assert np.array_equal(out, g['out0'])
    assert np.array_equal(info['negative_y_count'], g['negative_y_count'])
    # memory-light streaming variant used as bench cpu baseline
    tt, grid, where = orc.step_grid(g['tt'], 50)
    xT, neg = orc.heston_stream(
        dict(mu=.03, sigma=1., theta=.04, k=2., xi=.3), 100., .04, -.7, grid,
        101, np.random.default_rng(int(g['seed'])))
    assert np.array_equal(xT, g['out0'][-1])
    assert np.array_equal(neg, g['negative_y_count'])


@pytest.mark.parametrize('name,law', [
    ('seeded_merton', ('norm', dict(a=MERTON['a'], b=MERTON['b']))),
    ('seeded_kou', ('double_exp', dict(a=KOU['a'], b=KOU['b'], pa=KOU['pa']))),
])
def test_seeded_jumps(name, law):
    g = golden(name)
    out, info = orc.self_driven(
        'jumpdiff', dict(mu=.05, sigma=.2), 1., g['tt'], 80, 101,
        np.random.default_rng(int(g['seed'])), lam=15.,
        law=orc.jump_law(law[0], **law[1]))
    assert np.array_equal(out, g['out0'])
    assert np.array_equal(info['jump_count'], g['jump_count'])


def test_stats_against_reference():
    g = golden('stats_lognorm')
    x = g['x']
    assert np.array_equal(orc.pmean(x), g['pmean'])
    assert np.array_equal(orc.pvar(x), g['pvar'])
    assert np.array_equal(orc.pstd(x), g['pstd'])
    assert np.array_equal(orc.pvar(x, ddof=1), g['pvar1'])
    one = orc.moments_histogram(bins=25)
    one.update(x[-1])
    chunk = orc.moments_histogram(bins=25)
    for c in (x[-1][:, :1500], x[-1][:, 1500:2700], x[-1][:, 2700:]):
        chunk.update(c)
    for tag, m in (('one', one), ('chunk', chunk)):
        assert np.array_equal(m.mean(), g[tag + '_mean'])
        assert np.array_equal(m.var(), g[tag + '_var'])
        assert np.array_equal(m.std(), g[tag + '_std'])
        assert np.array_equal(m.stderr(), g[tag + '_stderr'])
        assert np.array_equal(m.skew(), g[tag + '_skew'])
        assert np.array_equal(m.kurtosis(), g[tag + '_kurt'])
        assert np.array_equal(np.stack(m.counts), g[tag + '_counts'])
        assert np.array_equal(np.stack(m.edges), g[tag + '_edges'])
        assert np.array_equal(np.array(m.outside), g[tag + '_outpaths'])


def test_known_answer_lognorm_exact():
    """The reference's own known-answer test (sdepy/tests/test_processes.py:
    681-708): Euler on log x is exact for constant parameters."""
    g = golden('known_lognorm_exact')
    t = g['t']
    tt, grid, where = orc.step_grid(t, None)
    out, _ = orc.euler_replay('lognorm', dict(mu=.05, sigma=.2), 1., grid,
                              where, g['dW'])
    exact = np.exp((.05 - .2*.2/2)*t[:, None] + .2*g['w'])
    assert np.allclose(out, exact, rtol=16*np.finfo(float).resolution)
    assert np.allclose(out, g['x'], rtol=1e-13)


TIME_AXIS = ('tmin', 'tmax', 'tsum', 'tmean', 'tvar', 'tstd', 'tcumsum', 'tder',
             'tint', 'tdiff', 'vmin', 'vmax', 'vsum', 'vmean', 'vvar', 'vstd')


def test_host_process_time_and_value_summaries_match_reference():
    """The host container's t* / v* methods (reference infrastructure.py:
    894-1122), bit for bit against the reference's own results."""
    import sdepy_b200 as m
    g = golden('stats_time_axis')
    p = m.process(t=g['t'], x=g['x'])
    for k in TIME_AXIS:
        r = getattr(p, k)()
        assert np.array_equal(np.asarray(r), g[k]), k
    assert np.array_equal(np.asarray(p.tvar(ddof=1)), g['tvar1'])
    assert np.array_equal(np.asarray(p.vstd(ddof=1)), g['vstd1'])
    q = p.tdiff(dt_exp=.5, fwd=False)
    assert np.array_equal(np.asarray(q), g['tdiff_half_bwd'])
    assert np.array_equal(q.t, g['tdiff_half_bwd_t'])
    assert np.array_equal(p.tmin().t, g['tmin_t'])


def test_host_process_container_matches_reference():
    """``p['t'|'p'|'v', ...]`` indexing, rebase, shapeas, piecewise and the
    timeline carried through ufuncs, against the reference's own results."""
    import sdepy_b200 as m
    keys = {
        't_int': ('t', 2), 't_last': ('t', -1), 't_slice': ('t', slice(1, 4)),
        'p_int': ('p', 3), 'p_slice': ('p', slice(0, 5, 2)), 'p_list': ('p', [1, 3]),
        'v_int': ('v', 1), 'v_two': ('v', 1, 2), 'v_tuple': ('v', (0, 1)),
        'v_mixed': ('v', slice(None), 0)}
    g = golden('process_container')
    p = m.process(g['t'], x=g['x'])
    for name, key in keys.items():
        q = p[key]
        assert isinstance(q, m.process), name
        assert np.array_equal(np.asarray(q), g[name]) and np.array_equal(q.t, g[name + '_t']), name
    assert not isinstance(p[2], m.process) and np.array_equal(p[1:3, 0], g['x'][1:3, 0])
    with pytest.raises(IndexError):
        p['w', 0]
    q = p.rebase((0.1, .55))
    assert np.array_equal(np.asarray(q), g['rebase']) and np.array_equal(q.t, g['rebase_t'])
    assert np.array_equal(np.asarray(p['v', 0].shapeas((4, 3))), g['shapeas'])
    with pytest.raises(ValueError):
        p.shapeas(())
    c = m.process(c=(1., 2., 3.))
    for q in (c.shapeas(p['v', 0])*p['v', 0], p['v', 0]*c.shapeas(p['v', 0])):
        assert isinstance(q, m.process)
        assert np.array_equal(np.asarray(q), g['const_times']) and np.array_equal(q.t, g['const_times_t'])
    a, b = p.pcopy(), p.tcopy()
    assert not np.shares_memory(a.t, p.t) and not np.shares_memory(a, p) and np.shares_memory(b, p)
    assert np.shares_memory(p.xcopy().t, p.t)
    for mode in ('mid', 'forward', 'backward'):
        q = m.piecewise(g['t'], v=g['x'][..., 0], mode=mode)
        assert np.array_equal(q(g['s']), g['piecewise_' + mode]), mode
    with pytest.raises(ValueError):
        m.piecewise(g['t'], v=g['x'][..., 0], mode='centre')


def _milstein_cases():
    g = golden('replay_milstein_plugin')

    def gbm(t, x, mu=.05, sigma=.4):
        return {'dt': mu*x, 'dw': sigma*x}

    def cev(t, x, mu=0., sigma=1.):
        return {'dt': mu*x, 'dw': sigma*x**1.5}
    yield ('gbm', gbm, dict(mu=.05, sigma=.4), 1., lambda t, x, mu, sigma: sigma, g)
    yield ('cev', cev, dict(mu=lambda t: .02 + .03*t, sigma=lambda t: .3 - .1*t), .8,
           lambda t, x, mu, sigma: 1.5*sigma*x**.5, g)


def test_milstein_pinned_to_the_reference_machinery_with_a_plugged_in_scheme():
    """The reference ships no Milstein; its documented ``method='<id>'`` hook
    (integration.py:675-685) runs one: tests/golden/make_milstein_plugin.py adds a
    ``milstein_next`` to a class made by the unmodified ``sdepy.integrate`` and
    records the result.  The oracle's restatement of the whole run (grid, left-point
    parameters, Euler part, correction term) must reproduce it bit for bit."""
    for name, f, par, x0, b_dx, g in _milstein_cases():
        grid, tt, dW, want = (g[name + '_' + k] for k in ('grid', 'tt', 'dW', 'x'))
        where = [int(np.flatnonzero(grid == t)[0]) for t in tt]
        got = orc.generic_replay(f, par, x0, grid, where, dW, scheme='milstein',
                                 diffusion_dx=b_dx)
        assert np.array_equal(got, want), name


def test_user_hooks_and_two_jump_terms_bit_exact():
    """The oracle's replay drivers with a user let / info_next and with both a
    'dn' and a 'dj' term, against the fixtures the reference produced for them
    (tests/golden/make_user_hooks.py, make_two_jump_terms.py)."""
    from tests.cases import two_jump_terms
    g = golden('replay_user_hooks')
    where = [int(np.searchsorted(g['a_grid'], t)) for t in g['a_tt']]
    xx, info = orc.generic_replay(
        lambda t, x, k=1., s=.4: {'dt': -k*x, 'dw': s}, dict(k=g['a_k']), .3, g['a_grid'], where,
        g['a_dW'], let=lambda x: x*x + 1.,
        info_next=lambda last, new: {'neg': (last < 0)*1, 'big': (new > .25)*1})
    assert np.array_equal(xx, g['a_out'])
    assert np.array_equal(info['neg'], g['a_neg']) and np.array_equal(info['big'], g['a_big'])
    assert int(info['neg'].sum()) == int(g['a_total_neg'])
    where = [int(np.searchsorted(g['b_grid'], t)) for t in g['b_tt']]
    xx, info = orc.system_replay(
        lambda t, x, y, a=.5: ({'dt': -a*x, 'dw': y}, {'dt': a*(1 - y), 'dw': .2*y}), 2, {},
        (1., .8), g['b_grid'], where, g['b_dW'], False, let=lambda x, y: x*y,
        info_next=lambda last, new: {'ylow': (last[1] < .9)*1})
    assert np.array_equal(xx, g['b_out']) and np.array_equal(info['ylow'], g['b_ylow'])
    g = golden('replay_two_jump_terms')
    where = [int(np.searchsorted(g['grid'], t)) for t in g['tt']]
    xx = orc.generic_replay(two_jump_terms, dict(a=g['p_a']), 1., g['grid'], where, g['dW'],
                            dN=g['dN'], dJ=g['dJ'])
    assert np.array_equal(xx, g['out'])
