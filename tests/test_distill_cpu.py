"""CPU tests of the distillation step's host logic (t2v_turbo_b200/distill.py) against tables and a tensor-level composition
produced by the UNMODIFIED reference (tests/golden/distill_tables.pt, oracle/make_goldens.py::gen_distill_tables): the DDIM
solver's tables, the boundary scalings, the guidance-scale embedding, and the folding of every per-sample affine combination
of train_t2v_turbo_v1_lora.py:1030-1039,1108-1181 into two coefficients per tensor (DistillStep.host_draws)."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tables():
    return torch.load(os.path.join(GOLD, "distill_tables.pt"))


def _step():
    from t2v_turbo_b200.distill import DistillStep
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    return DistillStep(None, None, T2VTurboScheduler(linear_start=0.00085, linear_end=0.012), num_ddim_timesteps=50, topk=20)


def test_ddim_solver_tables(tables):
    from t2v_turbo_b200.distill import DDIMSolver
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    s = DDIMSolver(T2VTurboScheduler(linear_start=0.00085, linear_end=0.012).alphas_cumprod.numpy(), ddim_timesteps=50)
    assert torch.equal(s.ddim_timesteps, tables["ddim_timesteps"])
    assert torch.equal(s.ddim_alpha_cumprods, tables["ddim_alpha_cumprods"])
    assert torch.equal(s.ddim_alpha_cumprods_prev, tables["ddim_alpha_cumprods_prev"])


def test_scalings_and_guidance_embedding(tables):
    from t2v_turbo_b200.distill import guidance_scale_embedding, scalings_for_boundary_conditions
    cs, co = scalings_for_boundary_conditions(tables["scal_t"], timestep_scaling=10.0)
    assert torch.equal(cs, tables["c_skip"]) and torch.equal(co, tables["c_out"])
    assert torch.equal(guidance_scale_embedding(tables["w"], embedding_dim=256), tables["w_emb"])


def test_coefficient_folding_reproduces_the_reference_composition(tables):
    """a[r] * x + b[r] * y with the host_draws coefficients == the reference's chain of get_predicted_original_sample /
    get_predicted_noise / CFG / ddim_step / boundary scalings, evaluated in fp64 on the same tensors."""
    c = tables["compose"]
    H = _step().host_draws(3, fixed=dict(index=c["index"], w=c["w"]))
    assert torch.equal(H["start_timesteps"], c["start"]) and torch.equal(H["timesteps"], c["tn"])

    def rows(a, x, b, y):
        return H[a].double().view(-1, 1, 1, 1, 1) * x + H[b].double().view(-1, 1, 1, 1, 1) * y
    z = rows("an_a", c["lat"], "an_b", c["noise"])
    model_pred = rows("k_z", z, "k_e", c["e_s"])
    eps_cfg = rows("cfg_c", c["e_c"], "cfg_u", c["e_u"])
    x_prev = rows("dd_x", rows("x0_z", z, "x0_e", eps_cfg), "dd_e", eps_cfg)
    target = rows("tg_x", x_prev, "tg_e", c["e_t"])
    for name, got in (("z", z), ("model_pred", model_pred), ("x_prev", x_prev), ("target", target)):
        err = ((got - c[name]).norm() / c[name].norm()).item()
        assert err < 2e-6, (name, err)          # the coefficients are stored in fp32


def test_host_draws_ranges():
    st = _step()
    g = torch.Generator().manual_seed(0)
    H = st.host_draws(64, generator=g)
    assert H["index"].min() >= 0 and H["index"].max() < 50
    assert ((H["w"] >= 5.0) & (H["w"] <= 15.0)).all()
    assert (H["timesteps"] == torch.clamp(H["start_timesteps"] - 20, min=0)).all()
    assert all(H[k].dtype == torch.float32 and H[k].shape == (64,) for k in st.COEFS)
    assert H["w_emb"].shape == (64, 256)


def test_ddim_reverse_step_coefficients(tables):
    """DDIM inversion (ode_solver/ddim_solver.py:89-97) as two per-sample coefficients == the reference's fp64 arithmetic, incl. the clip
    of the previous timestep at 0."""
    from t2v_turbo_b200.distill import DDIMSolver
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    r = tables["reverse"]
    s = DDIMSolver(T2VTurboScheduler(linear_start=0.00085, linear_end=0.012).alphas_cumprod.numpy(), ddim_timesteps=50)
    assert s.step_ratio == r["step_ratio"]
    ca, cb = s.reverse_coefs(r["ts"])
    x_t = ca.view(-1, 1, 1, 1, 1) * r["x_prev"] + cb.view(-1, 1, 1, 1, 1) * r["eps"]
    assert ((x_t - r["x_t"]).abs().max() / r["x_t"].abs().max()).item() < 1e-6       # the scheduler's abar table is fp32
