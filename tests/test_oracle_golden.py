"""CPU: pin the oracle (numpy restatement) to the golden fixtures produced by running the reference
(tests/golden/make_golden.py).  Tolerances are last-digit: the reference's own njit-vs-py bar is 1e-14..3e-14
(src/stochvolmodels/tests/test_logsv_characterization.py:403-404,439-441)."""
import numpy as np
import pytest

from conftest import chain_from_golden, load_golden
from oracle import mc, mgf


def test_time_grid_rule():
    for ttm, n, S, dt in load_golden("time_grid.npz")["cases"]:
        s2, dt2 = mc.set_time_grid(ttm, int(n))
        assert s2 == int(S) and dt2 == dt


def test_phi_grid_and_legacy_weights():
    g = load_golden("grids.npz")
    for name, vs, spot in (("mma", 0.2041241452319315, True), ("inv", 0.2041241452319315, False), ("dflt", 0.28, True)):
        phi = mgf.phi_grid(vs, spot)
        np.testing.assert_array_equal(phi, g[f"phi_{name}"])
        w = mgf.legacy_simpson_weights(phi)
        np.testing.assert_allclose(w, g[f"w_{name}"], rtol=1e-15, atol=0)
        h = phi[1].imag - phi[0].imag
        np.testing.assert_allclose(w[[0, 1, 2, -1]] / (h / 3), [1, 4, 2, 4], rtol=1e-14)   # even grid: last weight is 4
    assert mgf.logsv_vol_scaler(1.0, np.array([0.25])) == float(g["vol_scaler_q_025"])
    assert mgf.logsv_vol_scaler(0.8376, np.array([0.04289242541152263])) == float(g["vol_scaler_btc"])


def test_mlh_terms_and_rhs():
    g = load_golden("mlh.npz")
    theta, kappa1, kappa2, beta, volvol = g["params"]
    for k in range(int(g["ncases"])):
        order, spot, eta, pr, pi, sr, si = g[f"case{k}_in"]
        phi, psi = complex(pr, pi), complex(sr, si)
        M, L, H = mgf.logsv_mlh(theta, kappa1, kappa2, beta, volvol, phi, psi, bool(spot), int(order), eta)
        np.testing.assert_allclose(M, g[f"case{k}_M"], rtol=0, atol=5e-15)
        np.testing.assert_allclose(L, g[f"case{k}_L"], rtol=1e-15, atol=5e-15)
        np.testing.assert_allclose(H, g[f"case{k}_H"], rtol=1e-15, atol=5e-15)
        n = M.shape[0]
        A = (np.arange(1, n + 1) * (0.1 - 0.05j)).astype(np.complex128)
        np.testing.assert_allclose(mgf.rhs_dense(A, M, L, H), g[f"case{k}_rhs"], rtol=1e-14, atol=1e-14)
        fast = mgf.LogsvRhs(theta, kappa1, kappa2, beta, volvol, np.array([phi]), np.array([psi]), bool(spot), int(order), eta)
        np.testing.assert_allclose(fast(A[None, :])[0], g[f"case{k}_rhs"], rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("tag", ["g1_quickstart", "g2_inverse", "c3_5x21", "btc", "first_order", "backbone_inverse"])
def test_logsv_fourier_chain(tag):
    g = load_golden(f"logsv_fourier_{tag}.npz")
    strikes, types = chain_from_golden(g)
    prices, grids = mgf.logsv_chain_prices(g["params"], g["ttms"], g["forwards"], g["discfactors"], strikes, types,
                                           bool(g["is_spot"]), int(g["order"]), g["etas"], return_grids=True)
    np.testing.assert_array_equal(mgf.phi_grid(float(g["vol_scaler"]), bool(g["is_spot"])), g["phi"])
    for m in range(int(g["nslices"])):
        np.testing.assert_allclose(grids[m][0], g[f"a_t1_{m}"], rtol=0, atol=5e-12)
        np.testing.assert_allclose(grids[m][1], g[f"log_mgf_{m}"], rtol=0, atol=5e-12)
        np.testing.assert_allclose(prices[m], g[f"prices_{m}"], rtol=1e-11, atol=0)


def test_quickstart_value():
    g = load_golden("logsv_fourier_g1_quickstart.npz")
    assert abs(g["prices_0"][2] - 0.197331) < 5e-6 * 0.197331 + 1e-8      # examples/getting_started/quickstart.py:43
    assert abs(g["prices_1"][2] - 0.275202) < 5e-6 * 0.275202 + 1e-8      # :45


def test_rk45_statistics_match_survey_probe():
    g = load_golden("logsv_fourier_g1_quickstart.npz")
    sigma0, theta, k1, k2, beta, volvol = g["params"]
    rhs = mgf.LogsvRhs(theta, k1, k2, beta, volvol, g["phi"], np.zeros_like(g["phi"]), True, 2, 1.0)
    _, st = mgf.rk45_grid(rhs, np.zeros((1000, 5), dtype=complex), 0.25, return_stats=True)
    assert 5 <= st["nsteps"].min() and st["nsteps"].max() <= 60 and 40 <= st["nfev"].min() and st["nfev"].max() <= 500


@pytest.mark.parametrize("tag", ["g4", "c3_5x21", "btc"])
def test_heston_fourier_chain(tag):
    g = load_golden(f"heston_fourier_{tag}.npz")
    strikes, types = chain_from_golden(g)
    prices, grids = mgf.heston_chain_prices(g["params"], g["ttms"], g["forwards"], g["discfactors"], strikes, types, return_grids=True)
    for m in range(int(g["nslices"])):
        np.testing.assert_allclose(grids[m][0], g[f"log_mgf_{m}"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(grids[m][1], g[f"a_t1_{m}"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(grids[m][2], g[f"b_t1_{m}"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(prices[m], g[f"prices_{m}"], rtol=1e-11, atol=1e-14 * g["forwards"][m])   # deep-OTM prices ~1e-8 are F - K*capped cancellation


@pytest.mark.parametrize("tag", ["mma", "inv"])
def test_fourier_sum_on_lognormal_mgf(tag):
    g = load_golden(f"fourier_sum_{tag}.npz")
    p = mgf.vanilla_slice_prices(g["log_mgf"], g["phi"], float(g["forward"]), g["strikes"], g["types"], float(g["discfactor"]), bool(g["is_spot"]))
    np.testing.assert_allclose(p, g["prices"], rtol=1e-13)


def test_fourier_sum_error_conventions():
    g = load_golden("fourier_sum_mma.npz")
    with pytest.raises(ValueError, match="not implemented"):          # utils/mgf_pricer.py:206-212
        mgf.vanilla_slice_prices(g["log_mgf"], g["phi"], 1.5, g["strikes"][:1], np.array(["IC"]), 1.0, True)


def _fixed_randoms(g):
    rng = np.random.RandomState(int(g["seed"]))
    N = int(g["nb_path"])
    Z0s, Z1s = [], []
    for S in g["nsteps"]:                      # draw order: per slice W0 then W1 (pricers/logsv_pricer.py:1070-1071)
        Z0s.append(rng.normal(0, 1, size=(int(S), N)))
        Z1s.append(rng.normal(0, 1, size=(int(S), N)))
    return Z0s, Z1s


@pytest.mark.parametrize("tag", ["g5_c1", "inverse_eta", "btc_small", "qvar"])
def test_logsv_mc_fixed_randoms(tag):
    g = load_golden(f"logsv_mc_fixed_{tag}.npz")
    strikes, types = chain_from_golden(g)
    Z0s, Z1s = _fixed_randoms(g)
    np.testing.assert_array_equal(Z0s[0][0, :3], g["W0_head"])
    steps = mc.chain_steps(g["ttms"], int(g["n_per_year"]))
    assert [s for s, _ in steps] == list(g["nsteps"]) and np.all(np.array([d for _, d in steps]) == g["dts"])
    p, e, states = mc.logsv_mc_chain_fixed(g["params"], g["ttms"], g["forwards"], g["discfactors"], strikes, types, g["etas"],
                                           Z0s, Z1s, g["dts"], bool(g["is_spot"]), int(g["variable_type"]), True)
    for m in range(int(g["nslices"])):
        for a, name in zip(states[m], ("x", "sigma", "qvar")):
            np.testing.assert_allclose(a, g[f"{name}_{m}"], rtol=0, atol=3e-14)
        np.testing.assert_allclose(p[m], g[f"prices_{m}"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(e[m], g[f"stderr_{m}"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("tag", ["dflt", "floor"])
def test_heston_mc_stepper(tag):
    g = load_golden(f"heston_mc_fixed_{tag}.npz")
    N, S = int(g["nb_path"]), int(g["nsteps"])
    rng = np.random.RandomState(int(g["seed"]))
    Z0, Z1 = rng.normal(0, 1, size=(S, N)), rng.normal(0, 1, size=(S, N))
    v0, theta, kappa, rho, volvol = g["params"]
    x, v, q = mc.heston_step_fixed(np.zeros(N), v0 * np.ones(N), np.zeros(N), Z0, Z1, float(g["dt"]), theta, kappa, rho, volvol)
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(v, g["var"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(q, g["qvar"], rtol=0, atol=1e-14)
    assert v.min() >= 1e-4
    p, e = mc.mc_payoffs(x, q, float(g["ttm"]), float(g["forward"]), g["strikes"], g["types"], float(g["discfactor"]))
    np.testing.assert_allclose(p, g["prices"], rtol=1e-12)
    np.testing.assert_allclose(e, g["stderr"], rtol=1e-11)


def test_payoffs_all_codes_qvar_and_nan_paths():
    g = load_golden("payoffs.npz")
    kw = dict(ttm=float(g["ttm"]), forward=float(g["forward"]), discfactor=float(g["discfactor"]))
    for xs, suf in (("x", ""), ("x_nan", "_nan")):
        p, e = mc.mc_payoffs(g[xs], g["qvar"], strikes=g["strikes"], types=g["types"], **kw)
        np.testing.assert_allclose(p, g["prices" + suf], rtol=1e-13)
        np.testing.assert_allclose(e, g["stderr" + suf], rtol=1e-12)
        p, e = mc.mc_payoffs(g[xs], g["qvar"], strikes=g["qstrikes"], types=g["qtypes"], variable_type=mc.Q_VAR, **kw)
        np.testing.assert_allclose(p, g["qprices" + suf], rtol=1e-13)
        np.testing.assert_allclose(e, g["qstderr" + suf], rtol=1e-12)
    with pytest.raises(ValueError, match="payoff"):
        mc.mc_payoffs(g["x"], g["qvar"], strikes=np.ones(1), types=np.array(["BAD"]), **kw)
    with pytest.raises(NotImplementedError):
        mc.mc_payoffs(g["x"], g["qvar"], strikes=np.ones(1), types=np.array(["C"]), variable_type=mc.SIGMA, **kw)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def one(c, k):
        r = mc.philox4x32_10(*[np.array([v]) for v in c], k[0], k[1])
        return [int(v[0]) for v in r]
    assert one([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert one([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert one([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("gauss", ["f64", "f32"])
def test_device_normals_restatement_is_standard_normal(gauss):
    Z0, Z1 = mc.device_normals(10, np.arange(100000), 0, 3, gauss)
    z = np.concatenate([Z0.ravel(), Z1.ravel()])
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.std() - 1) < 4 / np.sqrt(2 * z.size)
    assert abs(np.mean(z ** 4) - 3) < 0.05 and abs(np.corrcoef(Z0.ravel(), Z1.ravel())[0, 1]) < 0.01
    # distinct slices / paths give distinct streams; same key replays
    A0, _ = mc.device_normals(10, np.arange(8), 1, 3, gauss)
    B0, _ = mc.device_normals(10, np.arange(8), 1, 3, gauss)
    np.testing.assert_array_equal(A0, B0)
    assert not np.any(A0 == Z0[:, :8])


@pytest.mark.parametrize("model", ["logsv", "heston"])
def test_c_port_matches_numpy_oracle(model):
    """the C port (CPU baseline / large-N checker) == numpy oracle fed with the restated device normals."""
    from oracle import cport
    K = np.array([0.8, 0.9, 1.0, 1.1, 1.2])
    ttms, fw, df = np.array([0.1, 0.3]), np.array([1.0, 1.02]), np.array([0.999, 0.99])
    strikes, N, npy, seed = [K, K], 3000, 252, 77
    steps = mc.chain_steps(ttms, npy)
    Z = [mc.device_normals(seed, np.arange(N), m, steps[m][0], "f64") for m in range(2)]
    dts = [d for _, d in steps]
    if model == "logsv":
        params = np.array([0.8376, 1.0413, 3.1844, 3.058, 0.1514, 1.8458])
        types = [np.array(["IP", "IP", "IC", "IC", "IC"]), np.array(["IP", "P", "C", "IC", "IC"])]
        etas = np.array([0.9, 1.1])
        p, e, st = mc.logsv_mc_chain_fixed(params, ttms, fw, df, strikes, types, etas, [z[0] for z in Z], [z[1] for z in Z], dts, False, 1, True)
        pc, ec, sc = cport.mc_chain("logsv", params, ttms, fw, df, etas, strikes, types, N, npy, False, 1, seed, "f64", return_states=True)
    else:
        params = np.array([0.04, 0.04, 4.0, -0.5, 0.4])
        types = [np.array(["P", "P", "C", "C", "C"])] * 2
        p, e, st = mc.heston_mc_chain_fixed(params, ttms, fw, df, strikes, types, [z[0] for z in Z], [z[1] for z in Z], dts, 1, True)
        pc, ec, sc = cport.mc_chain("heston", params, ttms, fw, df, None, strikes, types, N, npy, True, 1, seed, "f64", return_states=True)
    for a, b in zip(st[-1], sc):
        np.testing.assert_allclose(b, a, rtol=0, atol=2e-13)
    for m in range(2):
        np.testing.assert_allclose(pc[m], p[m], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(ec[m], e[m], rtol=1e-9, atol=1e-13)


# ---- quadratic-variance options, densities, digitals ----------------------------------------------------
@pytest.mark.parametrize("tag", ["mma", "inv"])
def test_logsv_qvar_fourier_chain(tag):
    g = load_golden(f"logsv_fourier_qvar_{tag}.npz")
    psi = mgf.psi_grid()
    assert psi.shape[0] == int(g["npsi"])
    np.testing.assert_array_equal(psi[:4], g["psi_head"])
    np.testing.assert_array_equal(psi[-2:], g["psi_tail"])
    K, T = g["strikes"], np.array(["C"] * len(g["strikes"]))
    prices, grids = mgf.logsv_qvar_chain_prices(g["params"], g["ttms"], g["discfactors"], [K, K], [T, T], bool(g["is_spot"]), 2, True)
    for m in range(2):
        np.testing.assert_allclose(grids[m][0][::40], g[f"a_sub_{m}"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(grids[m][1][::40], g[f"log_mgf_sub_{m}"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(prices[m], g[f"prices_{m}"], rtol=1e-10)
    with pytest.raises(ValueError, match="not implemented"):
        mgf.qvar_slice_prices(grids[0][1], psi, 0.25, K[:1], np.array(["P"]))


def test_logsv_pdfs_three_variables():
    g = load_golden("logsv_pdfs.npz")
    for vt, name in ((1, "LOG_RETURN"), (2, "Q_VAR"), (3, "SIGMA")):
        pdf = mgf.logsv_pdf(g["params"], float(g["ttm"]), g[f"grid_{name}"], vt)
        np.testing.assert_allclose(pdf, g[f"pdf_{name}"], rtol=1e-9, atol=1e-12)


def test_heston_qvar_fourier_chain():
    g = load_golden("heston_fourier_qvar.npz")
    K, T = g["strikes"], np.array(["C"] * len(g["strikes"]))
    prices = mgf.heston_qvar_chain_prices(g["params"], g["ttms"], g["discfactors"], [K, K], [T, T])
    for m in range(2):
        np.testing.assert_allclose(prices[m], g[f"prices_{m}"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("tag", ["neg", "pos"])
def test_digital_fourier_sums(tag):
    g = load_golden(f"fourier_digital_{tag}.npz")
    p = mgf.digital_slice_prices(g["log_mgf"], g["phi"], float(g["forward"]), g["strikes"], g["types"], float(g["discfactor"]))
    np.testing.assert_allclose(p, g["prices"], rtol=1e-12)


def test_logsv_vol_paths():
    g = load_golden("logsv_vol_paths.npz")
    S, dt = mc.set_time_grid(0.1, 360)
    W = np.sqrt(dt) * np.random.RandomState(5).normal(0, 1, size=(S, 500))
    for tag, spot in (("mma", True), ("inv", False)):
        sig = mc.logsv_vol_paths(*g["params"], W, dt, spot)
        np.testing.assert_allclose(sig, g[f"sigma_t_{tag}"], rtol=0, atol=1e-14)
    # LogSVPricer.simulate_vol_paths(ttm=0.02) passes ceil(360*0.02) = 8 as the PER-YEAR rate (logsv_pricer.py:574) => 1 step
    assert tuple(g["method_shape"]) == (2, 4)
    # the module function at 360 steps/yr gives 9 rows for ttm = 0.02 (reference tests/test_logsv_characterization.py:638-660)
    assert mc.set_time_grid(0.02, 360)[0] + 1 == 9


@pytest.mark.parametrize("name", ["h030_n3", "h045_n2", "h050_n1"])
def test_rough_mc_oracle_vs_reference_golden(name):
    """oracle/rough.py == the reference's rough-LogSV fixed-random chain pricer (terminal states per maturity, prices, 'std errors')"""
    from oracle import rough
    g = load_golden(f"rough_mc_{name}.npz")
    sigma0, theta, kappa1, kappa2, beta, volvol, H = g["params"]
    M, P = int(g["nslices"]), int(g["nb_path"])
    Z0, Z1, grids = rough.rough_randoms(g["ttms"], P, int(g["npy"]), int(g["seed"]))
    np.testing.assert_array_equal(Z0[:3, :5], g["Z0_head"])
    np.testing.assert_array_equal(Z1[:3, :5], g["Z1_head"])
    for m in range(M):
        np.testing.assert_array_equal(grids[m], g[f"grid_{m}"])
    prices, stds, states = rough.rough_chain_fixed(g["ttms"], g["forwards"], g["discfactors"], [g[f"strikes_{m}"] for m in range(M)],
                                                   [g[f"types_{m}"] for m in range(M)], Z0, Z1, sigma0, theta, kappa1, kappa2, beta, volvol,
                                                   g["weights"], g["nodes"], grids, return_states=True)
    for m in range(M):
        ls, vol, qv = states[m]
        np.testing.assert_allclose(ls, g[f"log_spot_{m}"], rtol=0, atol=2e-12)        # the reference kernels are fastmath=True
        np.testing.assert_allclose(vol, g[f"vol_{m}"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(qv, g[f"qv_{m}"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(prices[m], g[f"prices_{m}"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(stds[m], g[f"stds_{m}"], rtol=1e-9, atol=1e-14)


def test_rough_mc_oracle_vs_the_references_own_regression_fixture():
    """the reference's own known-answer vector for this path (tests/test_rough_logsv_pricer_regression.py, rtol 1e-7): BTC chain, H = 0.1"""
    from oracle import rough
    from stochvolmodels_b200 import get_btc_test_chain_data
    g = load_golden("rough_mc_reference_regression.npz")
    chain = get_btc_test_chain_data()
    sigma0, theta, kappa1, kappa2, beta, volvol, H = g["params"]
    Z0, Z1, grids = rough.rough_randoms(chain.ttms, int(g["nb_path"]), int(g["npy"]), int(g["seed"]))
    prices, _ = rough.rough_chain_fixed(chain.ttms, chain.forwards, chain.discfactors, chain.strikes_ttms, chain.optiontypes_ttms, Z0, Z1, sigma0, theta,
                                        kappa1, kappa2, beta, volvol, g["weights"], g["nodes"], grids)
    for m in range(int(g["nslices"])):
        np.testing.assert_allclose(prices[m], g[f"expected_prices_{m}"], rtol=1e-7, atol=0)


def test_hawkes_mc_oracle_vs_reference_golden():
    """oracle/hawkes.py == the reference's Hawkes jump-diffusion MC on the same legacy-generator draws (terminal states, chain prices)"""
    from oracle import hawkes
    g = load_golden("hawkes_mc.npz")
    for name in ("dflt", "drift"):
        params = dict(zip(hawkes.KEYS, g[f"{name}_params"]))
        rng = np.random.RandomState(int(g[f"{name}_seed"]))
        blk = hawkes.draw_inputs(rng, float(g[f"{name}_ttm"]), int(g[f"{name}_N"]), params["shift_p"], params["mean_p"], params["shift_m"], params["mean_m"])
        x, lp, lm = hawkes.step_fixed(g[f"{name}_x0"], g[f"{name}_lp0"], g[f"{name}_lm0"], *blk, **params)
        np.testing.assert_allclose(x, g[f"{name}_x"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(lp, g[f"{name}_lp"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(lm, g[f"{name}_lm"], rtol=1e-12, atol=1e-12)
    params = dict(zip(hawkes.KEYS, g["chain_params"]))
    M = g["chain_ttms"].shape[0]
    strikes = [g["chain_strikes"] * f for f in g["chain_forwards"]]
    prices, stds = hawkes.chain_prices(params, g["chain_ttms"], g["chain_forwards"], g["chain_discfactors"], strikes, [g["chain_types"]] * M,
                                       int(g["chain_N"]), rng=np.random.RandomState(int(g["chain_seed"])))
    for m in range(M):
        np.testing.assert_allclose(prices[m], g["chain_prices"][m], rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(stds[m], g["chain_stds"][m], rtol=1e-10, atol=1e-15)


@pytest.mark.parametrize("name", ["quick_first", "mild_second", "mild2_second_inverse"])
def test_semi_analytic_branch_oracle_vs_reference_golden(name):
    """the eig-free restatement of the semi-analytic branch == the reference's LAPACK route (every 8th grid point, three carried maturities)"""
    g = load_golden("logsv_analytic_branch.npz")
    sigma0, theta, k1, k2, beta, vv, order, spot = g[f"{name}_params"]
    phi = g[f"{name}_phi"][::8]
    n = mgf.expansion_n(int(order))
    a = np.zeros((phi.shape[0], n), dtype=np.complex128)
    t0 = 0.0
    for m, ttm in enumerate(g["ttms"]):
        a = mgf.logsv_analytic_a_grid(ttm - t0, phi, np.zeros_like(phi), a, theta, k1, k2, beta, vv, bool(spot), int(order))
        np.testing.assert_allclose(a, g[f"{name}_a_{m}"], rtol=1e-8, atol=1e-10)
        t0 = ttm


@pytest.mark.parametrize("name", ["quick_second", "btc_second_inverse_eta", "mild_first"])
def test_bdf_branch_oracle_vs_reference_golden(name):
    """SciPy BDF on the restated right-hand side / Jacobian == the reference's is_stiff_solver=True branch (every 8th grid point, 2 maturities)"""
    g = load_golden("logsv_bdf_branch.npz")
    sigma0, theta, k1, k2, beta, vv, order, spot, eta = g[f"{name}_params"]
    phi = g[f"{name}_phi"]
    a = np.zeros((phi.shape[0], mgf.expansion_n(int(order))), dtype=np.complex128)
    t0 = 0.0
    for m, ttm in enumerate(g["ttms"]):
        a = mgf.logsv_bdf_a_grid(ttm - t0, phi, np.zeros_like(phi), a, theta, k1, k2, beta, vv, bool(spot), int(order), eta)
        np.testing.assert_allclose(a, g[f"{name}_a_{m}"], rtol=1e-11, atol=1e-12)
        t0 = ttm


def test_hawkes_fourier_oracle_vs_reference_golden():
    """oracle/hawkes.py Fourier route (the Riccati system through the SciPy-RK45 clone) vs the reference's own outputs: ODE grids, log-MGF, chain
    prices, and the risk-kernel normalisers / forwards / prices"""
    from oracle import hawkes
    g = load_golden("hawkes_fourier.npz")
    K, T, ttms, fw, df = g["strikes"], g["types"], g["ttms"], g["forwards"], g["discfactors"]
    Ks, Ts = [K * f for f in fw], [T] * 3
    for name in ("dflt", "alt"):
        p = dict(zip(hawkes.KEYS, g[f"{name}_params"]))
        prices, grids, phi = hawkes.fourier_chain_prices(p, ttms, fw, df, Ks, Ts, return_grids=True)
        np.testing.assert_allclose(phi, g[f"{name}_phi"], rtol=0, atol=1e-13)
        for m in range(3):
            np.testing.assert_allclose(grids[m][0], g[f"{name}_a_{m}"], rtol=1e-11, atol=5e-12)
            np.testing.assert_allclose(grids[m][1], g[f"{name}_lm_{m}"], rtol=1e-11, atol=5e-12)
            np.testing.assert_allclose(prices[m], g[f"{name}_prices"][m], rtol=1e-11)
    p, gamma = dict(zip(hawkes.KEYS, g["dflt_params"])), float(g["gamma"])
    norm, gfw = hawkes.forwards_under_risk_kernel(p, gamma, ttms, fw)
    np.testing.assert_allclose(norm, g["gamma_normalizers"], rtol=1e-12)
    np.testing.assert_allclose(gfw, g["gamma_forwards"], rtol=1e-12)
    prices = hawkes.fourier_chain_prices(p, ttms, fw, df, Ks, Ts, gamma=gamma)
    for m in range(3):
        np.testing.assert_allclose(prices[m], g["gamma_prices"][m], rtol=1e-11)
