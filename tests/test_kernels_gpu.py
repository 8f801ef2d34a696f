"""GPU parity tests of the C-ABI kernels against the oracle (plain PyTorch fp32 / numpy).

Tolerances (stated per the north star): gathers / indexing / env transitions are BIT-EXACT;
floating-point kernels must match the fp32 oracle to max-relative error <= 1e-4 here
(the system-level bar is 1e-3 on logits and loss)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


@pytest.fixture(scope="module")
def setup():
    import clvln_b200
    from clvln_b200 import ops
    from clvln_b200.environ import make_world
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    world = make_world(n_scans=4, seed=3)
    store = ops.FeatureStore.from_world(world, dev)
    # kernel-level tests hold the weight-gradient library GEMMs to fp32 tolerances; the TF32 setting the
    # trainers use is covered by test_wgrad_tf32_tolerance and by the rollout-level gradient-cosine bars
    ops.WGRAD_TF32[0] = False
    yield world, store, ops, dev
    ops.WGRAD_TF32[0] = True


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def rand_state(world, B, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    vp = torch.randint(0, world.n_vp, (B,), generator=g, dtype=torch.int32)
    view = torch.randint(0, 36, (B,), generator=g, dtype=torch.int32)
    return vp.to(dev), view.to(dev)


def expected_pano(world, vp, view):
    from clvln_b200.environ.world import static_loc4
    loc4 = torch.from_numpy(static_loc4())
    img = world.table[vp.cpu().long()].float()
    ang = loc4[view.cpu().long()].repeat_interleave(32, dim=2)
    return torch.cat((img, ang), 2)


def expected_cand(world, vp, view, C):
    B = vp.shape[0]
    out = torch.zeros(B, C, 2176)
    for b in range(B):
        g, v = int(vp[b]), int(view[b])
        for j in range(int(world.n_cand[g])):
            out[b, j, :2048] = world.table[g, int(world.cand_view[g, j])].float()
            out[b, j, 2048:] = torch.from_numpy(world.cand_ang4[g, j, v % 12]).repeat_interleave(32)
    return out


def test_gather_pano_bit_exact(setup):
    world, store, ops, dev = setup
    vp, view = rand_state(world, 19, dev)
    out = ops.gather_pano(store, vp, view)
    assert torch.equal(out.cpu(), expected_pano(world, vp, view))


def test_gather_cand_bit_exact(setup):
    world, store, ops, dev = setup
    vp, view = rand_state(world, 23, dev, 1)
    out, lens = ops.gather_cand(store, vp, view)
    assert torch.equal(out.cpu(), expected_cand(world, vp, view, 16))
    assert torch.equal(lens.cpu().long(), torch.from_numpy(world.n_cand[vp.cpu().long().numpy()]).long() + 1)


@pytest.mark.parametrize("B", [13, 200])
@pytest.mark.parametrize("split", [1, 2, 4])
@pytest.mark.parametrize("drop_p", [0.0, 0.3])
def test_pano_attn_fwd_bwd(setup, split, drop_p, B):
    """B=13: one unit per CTA; B=200: persistent CTAs loop over several units (ring wrap-around)."""
    world, store, ops, dev = setup
    vp, view = rand_state(world, B, dev, 2)
    torch.manual_seed(split)
    q = (torch.randn(B, 2176, device=dev) * 0.05).requires_grad_(True)
    rng, off = ops.Rng(77, dev), 5
    out, attn = ops.pano_attn(store, vp, view, q, drop_p, rng, off, split)
    g_out = torch.randn_like(out)
    (dq,) = torch.autograd.grad(out, q, g_out)
    # oracle: SoftDotAttention maths (units.py:107-118) on the materialised, masked tensor
    img = expected_pano(world, vp, view).to(dev)
    if drop_p > 0:
        keep = ops.dropout_mask((B, 36, 2048), drop_p, rng, off).float()
        assert abs(keep.mean().item() - (1 - drop_p)) < 0.01
        img = torch.cat((img[..., :2048] * keep * (1.0 / (1.0 - drop_p)), img[..., 2048:]), -1)
    q2 = q.detach().clone().requires_grad_(True)
    logit = torch.einsum("bvf,bf->bv", img, q2)
    a = torch.softmax(logit, 1)
    ref = torch.einsum("bv,bvf->bf", a, img)
    (dq_ref,) = torch.autograd.grad(ref, q2, g_out)
    assert relerr(attn, a) < 1e-4
    assert relerr(out, ref) < 1e-4
    assert relerr(dq, dq_ref) < 1e-4


@pytest.mark.parametrize("drop_p", [0.0, 0.3])
def test_cand_logits_fwd_bwd(setup, drop_p):
    world, store, ops, dev = setup
    B = 21
    vp, view = rand_state(world, B, dev, 4)
    tgt = (torch.randn(B, 2176, device=dev) * 0.05).requires_grad_(True)
    bias = torch.randn(B, device=dev).requires_grad_(True)
    rng, off = ops.Rng(9, dev), 11
    logits = ops.cand_logits(store, vp, view, tgt, bias, drop_p, rng, off)
    cand = expected_cand(world, vp, view, 16).to(dev)
    if drop_p > 0:
        keep = ops.dropout_mask((B, 16, 2048), drop_p, rng, off).float()
        cand = torch.cat((cand[..., :2048] * keep * (1.0 / (1.0 - drop_p)), cand[..., 2048:]), -1)
    n = torch.from_numpy(world.n_cand[vp.cpu().long().numpy()]).to(dev)
    t2, b2 = tgt.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    ref = torch.einsum("bcf,bf->bc", cand, t2) + b2.unsqueeze(1)
    invalid = torch.arange(16, device=dev).unsqueeze(0) > n.unsqueeze(1)
    ref = ref.masked_fill(invalid, -math.inf)
    assert torch.equal(torch.isinf(logits), invalid)
    fin = ~invalid
    assert relerr(logits[fin], ref[fin]) < 1e-4
    g = torch.randn(B, 16, device=dev).masked_fill(invalid, 0.0)
    dt, db = torch.autograd.grad(logits, (tgt, bias), g)
    dt_ref, db_ref = torch.autograd.grad(ref.masked_fill(invalid, 0.0), (t2, b2), g)
    assert relerr(dt, dt_ref) < 1e-4 and relerr(db, db_ref) < 1e-4


@pytest.mark.parametrize("L,H", [(80, 512), (37, 256), (80, 1024)])
def test_ctx_attn_fwd_bwd(setup, L, H):
    from oracle import port_modules as P
    _, _, ops, dev = setup
    B = 9
    torch.manual_seed(L)
    ctx = torch.randn(B, L, H, device=dev).requires_grad_(True)
    tgt = (torch.randn(B, H, device=dev) * 0.1).requires_grad_(True)
    lengths = torch.randint(1, L + 1, (B,), device=dev, dtype=torch.int32)
    lengths[0] = L
    weighted, attn = ops.ctx_attn(ctx, tgt, lengths)
    mask = torch.arange(L, device=dev).unsqueeze(0) >= lengths.unsqueeze(1)
    c2, t2 = ctx.detach().clone().requires_grad_(True), tgt.detach().clone().requires_grad_(True)
    eye = torch.eye(H, device=dev)
    w_ref, a_ref = P.soft_dot_attention(t2, c2, eye, None, mask)     # linear_in = identity
    assert relerr(attn, a_ref) < 1e-4 and relerr(weighted, w_ref) < 1e-4
    gw, ga = torch.randn_like(weighted), torch.randn_like(attn)
    dc, dt = torch.autograd.grad((weighted, attn), (ctx, tgt), (gw, ga))
    dc_ref, dt_ref = torch.autograd.grad((w_ref, a_ref), (c2, t2), (gw, ga))
    assert relerr(dc, dc_ref) < 1e-4 and relerr(dt, dt_ref) < 1e-4


def test_lstm_pointwise(setup):
    _, _, ops, dev = setup
    B, H = 7, 512
    gates = torch.randn(B, 4 * H, device=dev, requires_grad=True)
    c0 = torch.randn(B, H, device=dev, requires_grad=True)
    h1, c1 = ops.lstm_pointwise(gates, c0)
    g2, c02 = gates.detach().clone().requires_grad_(True), c0.detach().clone().requires_grad_(True)
    i, f, g, o = g2.chunk(4, 1)
    c1r = torch.sigmoid(f) * c02 + torch.sigmoid(i) * torch.tanh(g)
    h1r = torch.sigmoid(o) * torch.tanh(c1r)
    assert relerr(h1, h1r) < 1e-5 and relerr(c1, c1r) < 1e-5
    gh, gc = torch.randn_like(h1), torch.randn_like(c1)
    dg, dc = torch.autograd.grad((h1, c1), (gates, c0), (gh, gc))
    dgr, dcr = torch.autograd.grad((h1r, c1r), (g2, c02), (gh, gc))
    assert relerr(dg, dgr) < 1e-4 and relerr(dc, dcr) < 1e-4


def test_policy_head(setup):
    import torch.nn.functional as F
    _, _, ops, dev = setup
    B = 33
    torch.manual_seed(0)
    logits = torch.randn(B, 16, device=dev) * 2
    nvalid = torch.randint(2, 17, (B,), device=dev)
    invalid = torch.arange(16, device=dev).unsqueeze(0) >= nvalid.unsqueeze(1)
    logits = logits.masked_fill(invalid, -math.inf).requires_grad_(True)
    target = (torch.rand(B, device=dev) * nvalid).long()
    target[::5] = -1
    # teacher + CE
    ce, logp, ent, act = ops.policy_head(logits, target, "teacher")
    l2 = logits.detach().clone().requires_grad_(True)
    ce_ref = F.cross_entropy(l2, target, ignore_index=-1, reduction="none")
    assert torch.equal(act.long(), target) and relerr(ce, ce_ref) < 1e-5
    cat = torch.distributions.Categorical(F.softmax(l2, 1))
    assert relerr(ent, cat.entropy()) < 1e-4
    w = torch.rand(B, device=dev)
    (d,) = torch.autograd.grad((ce * w).sum(), logits)
    (d_ref,) = torch.autograd.grad((ce_ref * w).sum(), l2)
    assert relerr(d, d_ref) < 1e-4
    # argmax
    _, _, _, a = ops.policy_head(logits, target, "argmax")
    assert torch.equal(a.long(), logits.max(1)[1])
    # sample: deterministic in (seed, offset), valid, log-prob/entropy gradients match autograd
    rng = ops.Rng(5, dev)
    ce, logp, ent, a = ops.policy_head(logits, target, "sample", rng, 9)
    _, _, _, a_again = ops.policy_head(logits, target, "sample", rng, 9)
    rng.off = 3
    rng.advance()                                                   # base += 3: a different stream
    _, _, _, a_other = ops.policy_head(logits, target, "sample", rng, 9)
    _, _, _, a_same = ops.policy_head(logits, target, "sample", rng, 6)
    assert not torch.equal(a, a_other) and torch.equal(a, a_same)
    rng = ops.Rng(5, dev)
    assert torch.equal(a, a_again)
    assert bool((a.long() < nvalid).all()) and bool((a >= 0).all())
    lp_ref = cat.log_prob(a.long())
    assert relerr(logp, lp_ref) < 1e-4
    g1, g2 = torch.randn(B, device=dev), torch.randn(B, device=dev)
    (d,) = torch.autograd.grad((logp * g1).sum() + (ent * g2).sum(), logits)
    (d_ref,) = torch.autograd.grad((lp_ref * g1).sum() + (cat.entropy() * g2).sum(), l2)
    assert relerr(d, d_ref) < 1e-4
    # sampling frequencies follow the softmax
    big = torch.tensor([[0.0, 1.0, 2.0, -1.0] + [-math.inf] * 12], device=dev).repeat(20000, 1)
    _, _, _, s = ops.policy_head(big, None, "sample", ops.Rng(1, dev), 2)
    freq = torch.bincount(s.long(), minlength=16).float() / 20000
    assert (freq - F.softmax(big[0], 0)).abs().max() < 0.015


def test_dropout(setup):
    _, _, ops, dev = setup
    x = torch.randn(1000, 37, device=dev, requires_grad=True)
    rng = ops.Rng(3, dev)
    rng.log = []
    y = ops.dropout(x, 0.5, rng, "t")
    assert rng.log == [("t", tuple(x.shape), 0.5, 1)]
    keep = ops.dropout_mask(x.shape, 0.5, rng, 1).float()
    assert torch.equal(y, x * keep * 2.0)
    assert abs(keep.mean().item() - 0.5) < 0.01
    (g,) = torch.autograd.grad(y.sum(), x)
    assert torch.equal(g, keep * 2.0)


@pytest.mark.parametrize("E,p", [(256, 0.5), (300, 0.5), (256, 0.0)])
def test_embed_dropout_fwd_bwd(setup, E, p):
    """Embedding + dropout in one kernel (units.py:48-52): forward bit-equal to embedding * the Philox mask of
    the dense tensor, backward equal to embedding_dense_backward of the masked gradient (padding row zero)."""
    _, _, ops, dev = setup
    torch.manual_seed(3)
    B, L, V = 16, 37, 992
    tok = torch.randint(0, V, (B, L), device=dev)
    tok[:, -5:] = 0                                                      # padding
    tok[0, :8] = 7                                                       # repeated entry
    w = torch.randn(V, E, device=dev, requires_grad=True)
    rng = ops.Rng(11, dev)
    rng.log = []
    y = ops.embed_dropout(tok, w, 0, p, rng)
    if p > 0:
        (tag, shape, pp, off), = rng.log
        assert tag == "enc_embed" and shape == (B, L, E)
        keep = ops.dropout_mask(shape, p, rng, off).float()
    else:
        assert rng.log == []
        keep = torch.ones(B, L, E, device=dev)
    w2 = w.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.embedding(tok, w2, padding_idx=0) * keep * (1.0 / (1.0 - p))
    assert torch.equal(y, ref)
    g = torch.randn_like(y)
    dw, = torch.autograd.grad(y, w, g)
    dref, = torch.autograd.grad(ref, w2, g)
    assert float(dw[0].abs().max()) == 0.0
    assert relerr(dw, dref) < 1e-5
    dw_again, = torch.autograd.grad(ops.embed_dropout(tok, w, 0, 0.0, rng), w, g)    # deterministic summation order
    dw_again2, = torch.autograd.grad(ops.embed_dropout(tok, w, 0, 0.0, rng), w, g)
    assert torch.equal(dw_again, dw_again2)


def test_dropout_unaligned_and_ragged(setup):
    """The vectorised dropout path and the scalar tail / unaligned fallback draw the same mask."""
    _, _, ops, dev = setup
    rng = ops.Rng(5, dev)
    base = torch.randn(4099, device=dev)
    for x in (base, base[1:], base[:4096]):
        y = ops._Dropout.apply(x.contiguous() if x.data_ptr() % 16 == 0 else x, 0.3, rng, 3)
        keep = ops.dropout_mask(x.shape, 0.3, rng, 3).float()
        sc = 1.0 / (1.0 - torch.tensor(0.3, dtype=torch.float32, device=dev))      # the kernel's fp32 1/(1-p)
        assert torch.equal(y, x * keep * sc)


def test_env_step_matches_world(setup):
    world, store, ops, dev = setup
    from clvln_b200.environ import make_items
    items = make_items(world, 24, seed=5)
    B = len(items)
    vp = torch.tensor([it["path_g"][0] for it in items], dtype=torch.int32, device=dev)
    goal = torch.tensor([it["path_g"][-1] for it in items], dtype=torch.int32, device=dev)
    view = torch.full((B,), 12, dtype=torch.int32, device=dev)
    ended = torch.zeros(B, dtype=torch.uint8, device=dev)
    teacher, dist = ops.env_observe(store, vp, ended, goal)
    cur = [it["path_g"][0] for it in items]
    for b in range(B):
        assert int(teacher[b]) == world.teacher_action(cur[b], int(goal[b]))
        assert float(dist[b]) == float(world.distance(cur[b], int(goal[b])))
    done = [False] * B
    n_active = torch.zeros(9, dtype=torch.int32, device=dev)
    for step in range(9):
        act = teacher.clone()
        vp0, view0 = vp.clone(), view.clone()
        vp2, view2, ended2, dist2, teacher, reward, mask = ops.env_step(store, vp, view, ended, dist, goal, act,
                                                                        n_active=n_active[step:step + 1])
        assert torch.equal(vp, vp0) and torch.equal(view, view0)       # inputs untouched
        for b in range(B):
            a = int(act[b])
            stop = done[b] or a < 0 or a >= int(world.n_cand[cur[b]])
            if not stop:
                exp_view = int(world.cand_view[cur[b], a])
                cur[b] = int(world.cand_vp[cur[b], a])
                assert int(view2[b]) == exp_view
            assert int(vp2[b]) == cur[b]
            assert float(mask[b]) == (0.0 if done[b] else 1.0)
            if not done[b] and stop:
                assert float(reward[b]) == 2.0                     # teacher path stops at the goal
            done[b] = done[b] or stop
            assert int(ended2[b]) == int(done[b])
            assert float(dist2[b]) == float(world.distance(cur[b], int(goal[b])))
            exp_t = -1 if done[b] else world.teacher_action(cur[b], int(goal[b]))
            assert int(teacher[b]) == exp_t
        assert int(n_active[step]) == sum(not d for d in done)
        vp, view, ended, dist = vp2, view2, ended2, dist2
    assert all(done)


def test_gather_action_feat(setup):
    world, store, ops, dev = setup
    B = 40
    vp, view = rand_state(world, B, dev, 8)
    g = torch.Generator().manual_seed(3)
    action = torch.randint(-1, 17, (B,), generator=g, dtype=torch.int32).to(dev)
    ended = (torch.rand(B, generator=g) < 0.2).to(torch.uint8).to(dev)
    out = ops.gather_action_feat(store, vp, view, action, ended)
    cands = expected_cand(world, vp, view, 16)
    for b in range(B):
        a, n = int(action[b]), int(world.n_cand[int(vp[b])])
        j = 0 if (int(ended[b]) or a < 0 or a >= n) else a
        assert torch.equal(out[b].cpu(), cands[b, j])


def test_a2c_loss(setup):
    """vln_a2c_fwd/bwd against the reference's numpy/torch recipe (envdrop.py:240-264)."""
    _, _, ops, dev = setup
    T, B, gamma = 9, 13, 0.9
    torch.manual_seed(1)
    reward = torch.randint(-2, 3, (T, B)).float().to(dev)
    ended_t = torch.rand(B) < 0.6
    stop_at = torch.randint(1, T + 1, (B,))
    mask = (torch.arange(T).unsqueeze(1) < torch.where(ended_t, stop_at, torch.full((B,), T)).unsqueeze(0)).float().to(dev)
    reward = reward * mask
    logp = (-torch.rand(T, B, device=dev)).requires_grad_(True)
    ent = torch.rand(T, B, device=dev).requires_grad_(True)
    value = torch.randn(T, B, device=dev).requires_grad_(True)
    last_value = torch.randn(B, device=dev)
    ended = ended_t.to(torch.uint8).to(dev)
    loss_b, stats = ops.a2c_loss(logp, ent, value, reward, mask, last_value, ended, gamma, 0.01)
    g = torch.rand(B, device=dev)
    d = torch.autograd.grad((loss_b * g).sum(), (logp, ent, value))
    # reference recipe
    l2, e2, v2 = (x.detach().clone().requires_grad_(True) for x in (logp, ent, value))
    disc = (~ended_t.numpy()) * last_value.cpu().numpy()
    ref = torch.zeros(B, device=dev)
    total, csq = 0, 0.0
    for t in range(T - 1, -1, -1):
        disc = disc * gamma + reward[t].cpu().numpy().astype(np.float64)
        m = mask[t].bool()
        r = torch.from_numpy(disc).to(dev)
        adv = (r - v2[t]).detach()
        cur = torch.zeros(B, device=dev)
        cur += (-l2[t] * adv * m)
        cur += (((r - v2[t]) ** 2) * m) * 0.5
        cur += (-0.01 * e2[t] * m)
        ref = ref + cur
        total += int(m.sum())
        csq += float((((r - v2[t]) ** 2) * m).sum())
    d_ref = torch.autograd.grad((ref * g).sum(), (l2, e2, v2))
    assert relerr(loss_b, ref.detach()) < 1e-5
    assert float(stats[0]) == total and abs(float(stats[1]) - csq) / csq < 1e-5
    for a, b in zip(d, d_ref):
        assert relerr(a, b) < 1e-5


@pytest.mark.parametrize("kind", [0, 1])
def test_fused_clip_optimizer(setup, kind):
    import ctypes as C
    from clvln_b200 import _lib
    from clvln_b200.ops import _ptr, _stream
    _, _, ops, dev = setup
    torch.manual_seed(kind)
    sizes = [5000, 12000, 700]
    params = [torch.randn(n, device=dev) for n in sizes]
    ref = [p.clone().requires_grad_(True) for p in params]
    flat = torch.cat(params).contiguous()
    s1, s2 = torch.zeros_like(flat), torch.zeros_like(flat)
    opt = (torch.optim.RMSprop if kind == 0 else torch.optim.Adam)(ref, lr=1e-2)
    off = (C.c_int64 * 4)(0, sizes[0], sizes[0] + sizes[1], sum(sizes))
    mx = (C.c_float * 3)(40.0, 1.5, 0.0)
    sq = torch.zeros(3, device=dev)
    for step in range(1, 4):
        grads = [torch.randn(n, device=dev) * (3.0 if i == 1 else 0.2) for i, n in enumerate(sizes)]
        for r, g in zip(ref, grads):
            r.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref[0]], 40.0)
        torch.nn.utils.clip_grad_norm_([ref[1]], 1.5)
        opt.step()
        gflat = torch.cat(grads).contiguous()
        _lib.check(_lib.lib().vln_grad_sqnorm(_ptr(gflat), off, 3, _ptr(sq), 1.0, _stream()))
        _lib.check(_lib.lib().vln_optim_step(_ptr(flat), _ptr(gflat), _ptr(s1), _ptr(s2), off, mx, 3, _ptr(sq), 1.0,
                                             kind, 1e-2, step, _stream()))
        assert relerr(flat, torch.cat([r.detach() for r in ref])) < 2e-5


@pytest.fixture
def lstm_variant(request):
    """Selects the recurrence kernels: 1 = tcgen05 (TMEM-resident W_hh), 0 = mma.sync (register-resident W_hh)."""
    from clvln_b200 import _lib
    _lib.check(_lib.lib().vln_lstm_set_variant(request.param))
    yield request.param
    _lib.check(_lib.lib().vln_lstm_set_variant(-1))


@pytest.mark.parametrize("lstm_variant", [1, 0], indirect=True, ids=["tcgen05", "mma"])
@pytest.mark.parametrize("H,E,B,L,ndir", [(256, 256, 64, 80, 2), (128, 300, 19, 33, 2), (256, 64, 5, 12, 1),
                                          (256, 128, 128, 40, 2), (128, 64, 150, 21, 2), (512, 256, 64, 40, 1),
                                          (512, 64, 21, 13, 2)])
def test_lstm_layer_matches_oracle(setup, lstm_variant, H, E, B, L, ndir, nb=None):
    """Persistent cluster LSTM (fwd + BPTT) vs the oracle's masked recurrence (== packed nn.LSTM); B = 128 / 150
    take the 24-rows-per-cluster instantiation of the tcgen05 kernels; H = 512 (Self-Monitor's uni-directional
    encoder) runs 16-CTA clusters with the lo half of W_hh in shared memory (tcgen05 path in both variants)."""
    from oracle import port_modules as P
    _, _, ops, dev = setup
    torch.manual_seed(H + B)
    x = torch.randn(B, L, E, device=dev)
    lengths = torch.randint(1, L + 1, (B,), device=dev).sort(descending=True)[0].to(torch.int32)
    lengths[0] = L
    ws = [(torch.randn(4 * H, E, device=dev) * 0.05, torch.randn(4 * H, H, device=dev) * 0.05,
           torch.randn(4 * H, device=dev) * 0.1) for _ in range(ndir)]
    mine = [[t.clone().requires_grad_(True) for t in w] for w in ws]
    ref = [[t.clone().requires_grad_(True) for t in w] for w in ws]
    xm, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    xproj = [ops.linear(xm.reshape(B * L, E), w[0], w[2]).view(B, L, 4 * H) for w in mine]
    out, h, c = ops.lstm_layer(xproj, [w[1] for w in mine], lengths)
    outs, hs, cs = [], [], []
    for k, w in enumerate(ref):
        o, hh, cc = P._lstm_direction(xr, lengths, w[0], w[1], w[2], torch.zeros_like(w[2]), reverse=bool(k))
        outs.append(o.transpose(0, 1)), hs.append(hh), cs.append(cc)
    out_r, h_r, c_r = torch.cat(outs, 2), torch.cat(hs, 1), torch.cat(cs, 1)
    assert relerr(out, out_r) < 1e-4 and relerr(h, h_r) < 1e-4 and relerr(c, c_r) < 1e-4
    assert bool((out[lengths.long().unsqueeze(1) <= torch.arange(L, device=dev).unsqueeze(0)] == 0).all())
    go, gh, gc = torch.randn_like(out), torch.randn_like(h), torch.randn_like(c)
    (out * go).sum().add((h * gh).sum()).add((c * gc).sum()).backward()
    (out_r * go).sum().add((h_r * gh).sum()).add((c_r * gc).sum()).backward()
    assert relerr(xm.grad, xr.grad) < 2e-4
    for wm, wr in zip(mine, ref):
        for a, b in zip(wm, wr):
            assert relerr(a.grad, b.grad) < 2e-4



@pytest.mark.parametrize("lstm_variant", [1], indirect=True, ids=["tcgen05"])
@pytest.mark.parametrize("nb", [16, 24, 32])
def test_lstm_tcgen05_rows_per_cluster(setup, lstm_variant, nb, monkeypatch):
    """Every rows-per-cluster instantiation (MMA N = 16 / 24 / 32) of the tcgen05 recurrence, ragged batch of 45."""
    monkeypatch.setenv("VLN_LSTM_NB", str(nb))
    test_lstm_layer_matches_oracle(setup, lstm_variant, 256, 128, 45, 30, 2)


@pytest.mark.parametrize("M,N,K", [(64, 2048, 2240), (64, 2176, 512), (7, 512, 1024), (100, 2048, 512), (64, 64, 128),
                                   (33, 1000, 192)])
def test_linear_tcgen05_bf16x3(setup, M, N, K):
    """tcgen05 bf16x3 skinny linear (fwd, input/weight/bias grads, accumulate) vs fp64 reference.
    Tolerance 2e-5 max-rel: three bf16 products recover ~16 mantissa bits."""
    _, _, ops, dev = setup
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev, requires_grad=True)
    w = (torch.randn(N, K, device=dev) * 0.05).requires_grad_(True)
    b = torch.randn(N, device=dev, requires_grad=True)
    acc = torch.randn(M, N, device=dev, requires_grad=True)
    assert ops.USE_TC_LINEAR[0]
    y = ops.linear(x, w, b, acc)
    ref = (x.double() @ w.double().t() + b.double() + acc.double())
    assert relerr(y.double(), ref) < 2e-5
    y2 = ops.linear(x, w)                                              # no bias, no accumulate
    assert relerr(y2.double(), x.double() @ w.double().t()) < 2e-5
    g = torch.randn_like(y)
    dx, dw, db, dacc = torch.autograd.grad(y, (x, w, b, acc), g)
    assert relerr(dx.double(), g.double() @ w.double()) < 2e-5
    assert relerr(dw.double(), g.double().t() @ x.double()) < 1e-5
    assert relerr(db, g.sum(0)) < 1e-5 and torch.equal(dacc, g)
    # the split cache follows in-place weight updates (autograd version counter)
    with torch.no_grad():
        w.mul_(0.5)
    assert relerr(ops.linear(x, w).double(), x.double() @ w.double().t()) < 2e-5


@pytest.mark.parametrize("M,N,K", [(5120, 1024, 256), (2368, 512, 512), (129, 256, 64), (700, 192, 320)])
def test_linear_tcgen05_tall(setup, M, N, K):
    """Tall activations (encoder input projection over B*L rows, batched critic): 128x128 output blocks,
    plain stores; forward + input gradient on the kernel vs fp64, ragged last row block included."""
    _, _, ops, dev = setup
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev, requires_grad=True)
    w = (torch.randn(N, K, device=dev) * 0.05).requires_grad_(True)
    b = torch.randn(N, device=dev, requires_grad=True)
    y = ops.linear(x, w, b)
    assert y.shape == (M, N)
    assert relerr(y.double(), x.double() @ w.double().t() + b.double()) < 2e-5
    g = torch.randn_like(y)
    dx, dw, db = torch.autograd.grad(y, (x, w, b), g)
    assert relerr(dx.double(), g.double() @ w.double()) < 2e-5
    assert relerr(dw.double(), g.double().t() @ x.double()) < 1e-5
    assert relerr(db, g.sum(0)) < 1e-4
    acc = torch.randn(M, N, device=dev)
    assert relerr(ops.linear(x, w, None, acc).double(), x.double() @ w.double().t() + acc.double()) < 2e-5


def test_wgrad_tf32_tolerance(setup):
    """dW = dY^T X over T*B stacked rows with TF32 inputs / fp32 accumulation: max-rel error vs fp64 stays
    within 2e-3 and the cosine within 1e-6 of 1 at the rollout's shapes."""
    _, _, ops, dev = setup
    torch.manual_seed(5)
    dy = torch.randn(2752, 2048, device=dev) * 0.1
    x = torch.randn(2752, 2752, device=dev)
    ref = dy.double().t() @ x.double()
    try:
        ops.WGRAD_TF32[0] = True
        got = ops.wgrad(dy, x).double()
    finally:
        ops.WGRAD_TF32[0] = False
    assert relerr(got, ref) < 2e-3
    cos = float((got * ref).sum() / (got.norm() * ref.norm()))
    assert cos > 1 - 1e-6
    assert relerr(ops.wgrad(dy, x).double(), ref) < 2e-5


# ---- fused decoder-step kernels (agent/fused.py) against their unfused counterparts ------------------
def test_feature_mask_bits_equal_dropout_mask(setup):
    """Packed keep-bits of several steps == vln_dropout_mask of each step's dense [B*36, 2048] tensor
    (bit-exact), in the byte order the panorama kernel consumes: byte (c//128)*128 + (c%32)*4 + (c%128)//32 = block c."""
    _, _, ops, dev = setup
    B, S, p = 5, 3, 0.3
    rng = ops.Rng(11, dev)
    bits = torch.empty((S, B * 36, 256), dtype=torch.uint8, device=dev)
    ops._call("vln_feature_mask_bits", ops._ptr(bits), B * 36, S, p, rng.ptr, 4, 7, ops._stream())
    c = torch.arange(256, device=dev)
    pos = (c // 128) * 128 + (c % 32) * 4 + (c % 128) // 32
    w = (1 << torch.arange(8, device=dev)).to(torch.int32)
    for s in range(S):
        keep = ops.dropout_mask((B * 36, 256, 8), p, rng, 4 + 7 * s).to(torch.int32)      # [rows, block, lane]
        want = (keep * w).sum(2).to(torch.uint8)                                          # byte of block c
        assert torch.equal(bits[s][:, pos], want)
    # sub-block generation (paired rollouts): rows [0, 2*36) of every step + rows [2*36, 5*36) of the first 2 steps
    sub = torch.zeros_like(bits)
    ops._call("vln_feature_mask_bits_ld", ops._ptr(sub), 2 * 36, B * 36, 0, S, p, rng.ptr, 4, 7, ops._stream())
    ops._call("vln_feature_mask_bits_ld", ops._ptr(sub), 3 * 36, B * 36, 2 * 36, 2, p, rng.ptr, 4, 7, ops._stream())
    assert torch.equal(sub[:2], bits[:2]) and torch.equal(sub[2, :72], bits[2, :72]) and int(sub[2, 72:].sum()) == 0


@pytest.mark.parametrize("mode", [0, 1])
def test_pano_attn_mask_bits_equal_inline_philox(setup, mode):
    """The kernel fed with pre-generated keep-bits computes exactly what it computes with inline Philox."""
    world, store, ops, dev = setup
    B, p = 21, 0.3
    vp, view = rand_state(world, B, dev, 5)
    rng = ops.Rng(3, dev)
    vec = torch.randn(B, 2176, device=dev) * 0.05
    attn = torch.softmax(torch.randn(B, 36, device=dev), 1)
    bits = torch.empty((B * 36, 256), dtype=torch.uint8, device=dev)
    ops._call("vln_feature_mask_bits", ops._ptr(bits), B * 36, 1, p, rng.ptr, 9, 0, ops._stream())
    outs = []
    for mb in (None, bits):
        a = attn.clone()
        out = torch.empty(B, 2176, device=dev)
        ops._call("vln_pano_attn_ld", store.handle, ops._ptr(vp), ops._ptr(view), ops._ptr(store.loc4), ops._ptr(vec), 2176,
                  ops._ptr(a), None, 2176, ops._ptr(out), 2176, B, mode, p, rng.ptr, 9, ops._ptr(mb), 1, ops._stream())
        outs.append((out, a))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("B", [21, 1500])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_pano_attn_streaming_kernel_matches_cluster_kernel(setup, mode, B, p, monkeypatch):
    """The two kernels behind vln_pano_attn_ld (variant 2 = 2-CTA cluster, 4 = streaming, single-pass online
    softmax) agree to fp32 rounding, with and without pre-generated keep-bits, into strided output rows.
    B=1500 > resident CTAs: every CTA loops over several episodes (ring wrap-around across episodes)."""
    world, store, ops, dev = setup
    vp, view = rand_state(world, B, dev, 6)
    rng = ops.Rng(5, dev)
    torch.manual_seed(mode)
    ld = 2176 + 64
    vec = torch.randn(B, ld, device=dev) * 0.05
    attn = torch.softmax(torch.randn(B, 36, device=dev) * 2, 1)
    bits = None
    if p > 0:
        bits = torch.empty((B * 36, 256), dtype=torch.uint8, device=dev)
        ops._call("vln_feature_mask_bits", ops._ptr(bits), B * 36, 1, p, rng.ptr, 9, 0, ops._stream())
    outs = []
    for variant in (2, 4):
        a = attn.clone()
        out = torch.full((B, ld), 7.0, device=dev)
        ops._call("vln_pano_attn_ld", store.handle, ops._ptr(vp), ops._ptr(view), ops._ptr(store.loc4), ops._ptr(vec), ld,
                  ops._ptr(a), None, ld, ops._ptr(out), ld, B, mode, p, rng.ptr, 9, ops._ptr(bits), variant, ops._stream())
        outs.append((out, a))
    assert torch.equal(outs[1][0][:, 2176:], torch.full((B, 64), 7.0, device=dev))      # nothing past column 2176
    assert relerr(outs[1][0][:, :2176], outs[0][0][:, :2176]) < 2e-5
    assert relerr(outs[1][0][:, 2048:2176], outs[0][0][:, 2048:2176]) < 2e-5            # angle columns on their own scale
    assert relerr(outs[1][1], outs[0][1]) < 2e-5
    # per-episode scale as well: a wrong episode hides in a global max
    num = (outs[1][0][:, :2176] - outs[0][0][:, :2176]).abs().amax(1)
    assert (num / outs[0][0][:, :2176].abs().amax(1)).max().item() < 1e-4


def test_policy_env_act_equals_separate_kernels(setup):
    """vln_policy_env_act_fwd == vln_policy_fwd + vln_env_step + vln_envdrop_act_fwd (bit-exact state, actions,
    rewards; identical floating-point results)."""
    world, store, ops, dev = setup
    B, E, p = 37, 64, 0.5
    torch.manual_seed(2)
    vp, view = rand_state(world, B, dev, 9)
    goal = torch.randint(0, world.n_vp, (B,), dtype=torch.int32).to(dev)
    # goals must lie in the same scan as the viewpoint for the distance tables: reuse the viewpoint's own scan
    goal = vp.clone()
    ended = (torch.rand(B, device=dev) < 0.2).to(torch.uint8)
    teacher, dist = ops.env_observe(store, vp, ended, goal)
    n = store.n_cand[vp.long()]
    logits = torch.randn(B, 16, device=dev)
    logits[torch.arange(16, device=dev).unsqueeze(0) > n.unsqueeze(1)] = float("-inf")
    rng = ops.Rng(5, dev)
    w_act, b_act = torch.randn(E, 128, device=dev) * 0.1, torch.randn(E, device=dev) * 0.1
    wg = w_act.view(E, 4, 32).sum(2).contiguous()
    for fb in (0, 1, 2):
        ce, logp, ent, action = ops.policy_head(logits, teacher, fb, rng, 3)
        vp2, view2, ended2, dist2, teacher2, reward, mask = ops.env_step(store, vp, view, ended, dist, goal, action)
        act = torch.empty(B, E, device=dev)
        xh = torch.zeros(B, 96, device=dev)
        ops._call("vln_envdrop_act_fwd", ops._ptr(view2), ops._ptr(store.pose4), ops._ptr(wg), ops._ptr(b_act), ops._ptr(act),
                  ops._ptr(xh), 96, B, E, p, rng.ptr, 4, ops._stream())
        o = dict(ce=torch.empty(B, device=dev), action=torch.empty(B, dtype=torch.int32, device=dev),
                 logp=torch.empty(B, device=dev), ent=torch.empty(B, device=dev), probs=torch.empty(B, 16, device=dev),
                 vp=torch.empty_like(vp), view=torch.empty_like(view), ended=torch.empty_like(ended),
                 dist=torch.empty_like(dist), teacher=torch.empty_like(teacher), reward=torch.empty(B, device=dev),
                 mask=torch.empty(B, device=dev), n_active=torch.zeros(1, dtype=torch.int32, device=dev),
                 act=torch.empty(B, E, device=dev), xh=torch.zeros(B, 96, device=dev))
        P = ops._ptr
        ops._call("vln_policy_env_act_fwd", P(logits), P(teacher), fb, rng.ptr, 3, P(o["ce"]), P(o["action"]), P(o["logp"]),
                  P(o["ent"]), P(o["probs"]), P(vp), P(view), P(ended), P(dist), P(goal), P(store.cand_vp), P(store.cand_view),
                  P(store.n_cand), P(store.next_hop), P(store.dist), P(store.sq_off), P(store.vp_local), P(o["vp"]),
                  P(o["view"]), P(o["ended"]), P(o["dist"]), P(o["teacher"]), P(o["reward"]), P(o["mask"]), P(o["n_active"]),
                  P(store.pose4), P(wg), P(b_act), P(o["act"]), P(o["xh"]), 96, E, p, 4, B, ops._stream())
        for a, b in ((ce, o["ce"]), (logp, o["logp"]), (ent, o["ent"]), (action, o["action"]), (vp2, o["vp"]),
                     (view2, o["view"]), (ended2, o["ended"]), (dist2, o["dist"]), (teacher2, o["teacher"]),
                     (reward, o["reward"]), (mask, o["mask"]), (act, o["act"]), (xh, o["xh"])):
            assert torch.equal(a, b)
        assert int(o["n_active"]) == int((ended2 == 0).sum())
        # the action embedding itself: tanh(W_a angle128(view') + b_a), nn.Linear on the 128-wide feature
        ref = torch.tanh(store.pose128[view2.long()] @ w_act.t() + b_act)
        assert relerr(act, ref) < 1e-5


@pytest.mark.parametrize("pf", [0.0, 0.3])
def test_fused_step_tail_equals_separate_kernels(setup, pf):
    """vln_cand_policy_env_act_fwd == vln_cand_logits_fwd followed by vln_policy_env_act_fwd, bit for bit (logits,
    probabilities, actions, new state, teacher slot, rewards, action embedding), incl. per-row teacher forcing."""
    world, store, ops, dev = setup
    B, E, p = 41, 64, 0.5
    torch.manual_seed(3)
    vp, view = rand_state(world, B, dev, 12)
    goal = vp.clone()
    goal[::3] = store.cand_vp[vp[::3].long(), 0]                  # a neighbour as the goal: non-trivial teacher slots
    ended = (torch.rand(B, device=dev) < 0.2).to(torch.uint8)
    teacher, dist = ops.env_observe(store, vp, ended, goal)
    tgt = torch.randn(B, 2176, device=dev) * 0.05
    rng = ops.Rng(5, dev)
    w_act, b_act = torch.randn(E, 128, device=dev) * 0.1, torch.randn(E, device=dev) * 0.1
    wg = w_act.view(E, 4, 32).sum(2).contiguous()
    P = ops._ptr

    def outputs():
        return dict(logits=torch.empty(B, 16, device=dev), ce=torch.empty(B, device=dev),
                    action=torch.empty(B, dtype=torch.int32, device=dev), logp=torch.empty(B, device=dev),
                    ent=torch.empty(B, device=dev), probs=torch.empty(B, 16, device=dev), vp=torch.empty_like(vp),
                    view=torch.empty_like(view), ended=torch.empty_like(ended), dist=torch.empty_like(dist),
                    teacher=torch.empty_like(teacher), reward=torch.empty(B, device=dev), mask=torch.empty(B, device=dev),
                    n_active=torch.zeros(1, dtype=torch.int32, device=dev), act=torch.empty(B, E, device=dev),
                    xh=torch.zeros(B, 96, device=dev))
    for fb in (0, 1, 2, 2 | ((B // 2 + 1) << 8)):
        a, o = outputs(), outputs()
        ops._call("vln_cand_logits_fwd", store.handle, P(vp), P(view), P(store.cand_view), P(store.cand_ang4), P(store.n_cand),
                  P(tgt), None, P(a["logits"]), B, pf, rng.ptr, 7, ops._stream())
        ops._call("vln_policy_env_act_fwd", P(a["logits"]), P(teacher), fb, rng.ptr, 3, P(a["ce"]), P(a["action"]), P(a["logp"]),
                  P(a["ent"]), P(a["probs"]), P(vp), P(view), P(ended), P(dist), P(goal), P(store.cand_vp), P(store.cand_view),
                  P(store.n_cand), P(store.next_hop), P(store.dist), P(store.sq_off), P(store.vp_local), P(a["vp"]),
                  P(a["view"]), P(a["ended"]), P(a["dist"]), P(a["teacher"]), P(a["reward"]), P(a["mask"]), P(a["n_active"]),
                  P(store.pose4), P(wg), P(b_act), P(a["act"]), P(a["xh"]), 96, E, p, 4, B, ops._stream())
        ops._call("vln_cand_policy_env_act_fwd", store.handle, P(vp), P(view), P(store.cand_ang4), P(tgt), P(o["logits"]), pf, 7,
                  P(teacher), fb, rng.ptr, 3, P(o["ce"]), P(o["action"]), P(o["logp"]), P(o["ent"]), P(o["probs"]),
                  P(ended), P(dist), P(goal), P(store.cand_vp), P(store.cand_view), P(store.n_cand), P(store.next_hop),
                  P(store.dist), P(store.sq_off), P(store.vp_local), P(o["vp"]), P(o["view"]), P(o["ended"]), P(o["dist"]),
                  P(o["teacher"]), P(o["reward"]), P(o["mask"]), P(o["n_active"]), P(store.pose4), P(wg), P(b_act), P(o["act"]),
                  P(o["xh"]), 96, E, p, 4, B, ops._stream())
        for k in a:
            assert torch.equal(a[k], o[k]), (fb, k)
        assert bool((a["teacher"][a["ended"] == 0] >= 0).all())


def test_cand_bwd_policy_stacked_steps(setup):
    """vln_cand_logits_bwd_policy over [n_steps, B] stacked rows == vln_policy_bwd + vln_cand_logits_bwd per step."""
    world, store, ops, dev = setup
    B, S, p = 11, 3, 0.3
    rng = ops.Rng(8, dev)
    g = torch.Generator().manual_seed(4)
    vp = torch.randint(0, world.n_vp, (S, B), generator=g, dtype=torch.int32).to(dev)
    view = torch.randint(0, 36, (S, B), generator=g, dtype=torch.int32).to(dev)
    probs = torch.softmax(torch.randn(S, B, 16, device=dev), 2)
    n = store.n_cand[vp.long()]
    probs = probs * (torch.arange(16, device=dev).view(1, 1, 16) <= n.unsqueeze(2))
    probs = probs / probs.sum(2, keepdim=True)
    ent = -(probs * torch.log(probs.clamp_min(1e-30))).sum(2)
    target = torch.randint(-1, 2, (S, B), dtype=torch.int32).to(dev).clamp_max(0) * 0 + torch.minimum(
        torch.randint(0, 16, (S, B), dtype=torch.int32).to(dev), n)
    action = torch.minimum(torch.randint(0, 16, (S, B), dtype=torch.int32).to(dev), n)
    target[0, 0], action[1, 1] = -1, -1
    g_ce, g_lp, g_en = (torch.randn(S, B, device=dev) for _ in range(3))
    P = ops._ptr
    d_all = torch.empty(S, B, 2176, device=dev)
    ops._call("vln_cand_logits_bwd_policy", store.handle, P(vp), P(view), P(store.cand_view), P(store.cand_ang4), P(store.n_cand),
              P(probs), P(target), P(action), P(ent), P(g_ce), P(g_lp), P(g_en), P(d_all), B, S, p, rng.ptr, 20, 5, ops._stream())
    for s in range(S):
        dl = torch.empty(B, 16, device=dev)
        ops._call("vln_policy_bwd", P(probs[s]), P(target[s]), P(action[s]), P(ent[s]), P(g_ce[s]), P(g_lp[s]), P(g_en[s]),
                  P(dl), B, ops._stream())
        d = torch.empty(B, 2176, device=dev)
        ops._call("vln_cand_logits_bwd", store.handle, P(vp[s]), P(view[s]), P(store.cand_view), P(store.cand_ang4),
                  P(store.n_cand), P(dl), P(d), None, B, p, rng.ptr, 20 + 5 * s, ops._stream())
        assert torch.equal(d, d_all[s])


def test_step_glue_kernels(setup):
    """state / LSTM-pointwise / action-embedding glue kernels against torch on the masks vln_dropout_mask yields."""
    _, _, ops, dev = setup
    B, H, p = 9, 512, 0.5
    torch.manual_seed(6)
    rng = ops.Rng(21, dev)
    P = ops._ptr
    sc = 1.0 / (1.0 - p)
    keep = lambda shape, off: ops.dropout_mask(shape, p, rng, off).float() * sc
    # state_fwd / state_bwd
    src = torch.randn(B, H, device=dev)
    xh = torch.zeros(B, H + 40, device=dev)
    hq, hc = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
    ops._call("vln_envdrop_state_fwd", P(src), 1, ops.C.c_void_p(xh.data_ptr() + 4 * 40), H + 40, P(hq), P(hc), B, H, p,
              rng.ptr, 2, 3, ops._stream())
    ht = torch.tanh(src)
    assert relerr(xh[:, 40:], ht) < 1e-6 and bool((xh[:, :40] == 0).all())
    assert relerr(hq, ht * keep((B, H), 2)) < 1e-6 and relerr(hc, ht * keep((B, H), 3)) < 1e-6
    d_hc, d_hq, d_x = (torch.randn(B, H, device=dev) for _ in range(3))
    d_src = torch.empty(B, H, device=dev)
    ops._call("vln_envdrop_state_bwd", P(d_hc), P(d_x), H, P(d_hq), ops.C.c_void_p(xh.data_ptr() + 4 * 40), H + 40, 1, P(d_src),
              B, H, p, rng.ptr, 2, 3, ops._stream())
    ref = (d_hc * keep((B, H), 3) + d_x + d_hq * keep((B, H), 2)) * (1 - ht * ht)
    assert relerr(d_src, ref) < 1e-5
    # LSTM pointwise + dropout, forward and backward, against the unfused kernels
    gates, c0 = torch.randn(B, 4 * H, device=dev), torch.randn(B, H, device=dev)
    h1, c1 = ops.lstm_pointwise(gates, c0)
    o = [torch.empty(B, H, device=dev) for _ in range(2)] + [torch.empty(B, 4 * H, device=dev), torch.zeros(B, 2 * H, device=dev)]
    ops._call("vln_lstm_pointwise_drop_fwd", P(gates), P(c0), P(o[0]), P(o[1]), P(o[2]), ops.C.c_void_p(o[3].data_ptr() + 4 * H),
              2 * H, B, H, p, rng.ptr, 6, ops._stream())
    assert relerr(o[0], h1) < 1e-6 and relerr(o[1], c1) < 1e-6
    assert relerr(o[3][:, H:], h1 * keep((B, H), 6)) < 1e-6 and bool((o[3][:, :H] == 0).all())
    d_drop, d_extra, d_c1 = torch.randn(B, 2 * H, device=dev), torch.randn(B, H, device=dev), torch.randn(B, H, device=dev)
    dg, dc0 = torch.empty(B, 4 * H, device=dev), torch.empty(B, H, device=dev)
    ops._call("vln_lstm_pointwise_drop_bwd", P(o[2]), P(c0), P(o[1]), ops.C.c_void_p(d_drop.data_ptr() + 4 * H), 2 * H, P(d_extra),
              P(d_c1), P(dg), P(dc0), B, H, p, rng.ptr, 6, ops._stream())
    g2, c2 = gates.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    hh, cc = ops.lstm_pointwise(g2, c2)
    ((hh * (d_drop[:, H:] * keep((B, H), 6) + d_extra)).sum() + (cc * d_c1).sum()).backward()
    assert relerr(dg, g2.grad) < 1e-5 and relerr(dc0, c2.grad) < 1e-5
    # batched action-embedding backward
    S, E = 3, 64
    act = torch.tanh(torch.randn(S, B, E, device=dev))
    d_xh = torch.randn(S, B, 96, device=dev)
    d_pre = torch.empty(S, B, E, device=dev)
    ops._call("vln_envdrop_act_bwd", P(d_xh), 96, P(act), P(d_pre), B, E, S, p, rng.ptr, 30, 4, ops._stream())
    for s in range(S):
        assert relerr(d_pre[s], d_xh[s, :, :E] * keep((B, E), 30 + 4 * s) * (1 - act[s] ** 2)) < 1e-5


def test_linear_pair_equals_two_launches(setup):
    _, _, ops, dev = setup
    torch.manual_seed(9)
    M, N, K = 64, 2176, 512
    ws = [torch.randn(N, K, device=dev) * 0.05 for _ in range(2)]
    xs = [torch.randn(M, K, device=dev) for _ in range(2)]
    sw = [ops._SplitWeight(w).fresh(w) for w in ws]
    ys = [torch.zeros(M, N, device=dev) for _ in range(2)]
    P = ops._ptr
    ops._call("vln_linear_bf16x3_pair", P(sw[0].hi), P(sw[0].lo), P(xs[0]), P(ys[0]), P(sw[1].hi), P(sw[1].lo), P(xs[1]), P(ys[1]),
              N, K, K, M, N, ops._stream())
    for w, x, y in zip(ws, xs, ys):
        ref = x.double() @ w.double().t()
        assert relerr(y.double(), ref) < 2e-5


@pytest.mark.parametrize("p", [0.0, 0.5])
@pytest.mark.parametrize("B,L", [(64, 80), (5, 33)])
def test_ctx_step_equals_separate_kernels(setup, B, L, p):
    """vln_envdrop_ctx_step_fwd/bwd (csrc/ctx_step.cu: LSTM pointwise + dropout + text attention in one launch, the
    query projection folded into the context as CW = ctx W_in) against the three launches it replaces
    (vln_lstm_pointwise_drop -> W_in GEMM -> vln_ctx_attn) in fp64-checked torch on the same inputs and masks."""
    import ctypes as C
    world, store, ops, dev = setup
    H = 512
    g = torch.Generator().manual_seed(B * 131 + L)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    gates, c0 = r(B, 4 * H), r(B, H)
    ctx = r(B, L, H) * 0.3
    w_in = r(H, H) * 0.05
    lengths = torch.randint(1, L + 1, (B,), generator=g).to(torch.int32).to(dev)
    lengths[0] = L
    cw = (ctx.double() @ w_in.double()).float().contiguous()
    rng = ops.Rng(11, dev)
    off = 5
    h1, c1, acts = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev), torch.empty(B, 4 * H, device=dev)
    wh = torch.zeros(B, 2 * H, device=dev)
    attn = torch.full((B, L), -1.0, device=dev)
    for ready in (0, 1):
        ops._call("vln_envdrop_ctx_step_fwd", ops._ptr(gates), ops._ptr(c0), ops._ptr(h1), ops._ptr(c1), ops._ptr(acts),
                  ops._ptr(wh), 2 * H, ops._ptr(ctx), ops._ptr(cw), ops._ptr(lengths), ops._ptr(attn), B, L, H, p, rng.ptr,
                  off, ready, ops._stream())
    # reference: the oracle's formulas in float64
    i, f, gg, o = gates.double().chunk(4, 1)
    i, f, gg, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)
    c_ref = f * c0.double() + i * gg
    h_ref = o * torch.tanh(c_ref)
    keep = ops.dropout_mask((B, H), p, rng, off).double() / (1.0 - p) if p > 0 else torch.ones(B, H, device=dev).double()
    hd = h_ref * keep
    tq = hd @ w_in.double().t()
    logit = torch.einsum("blh,bh->bl", ctx.double(), tq)
    mask = torch.arange(L, device=dev).unsqueeze(0) >= lengths.unsqueeze(1)
    a_ref = torch.softmax(logit.masked_fill(mask, float("-inf")), 1)
    w_ref = torch.einsum("bl,blh->bh", a_ref, ctx.double())
    assert relerr(h1.double(), h_ref) < 1e-5 and relerr(c1.double(), c_ref) < 1e-5
    assert relerr(acts.double(), torch.cat((i, f, gg, o), 1)) < 1e-5
    assert relerr(wh[:, H:].double(), hd) < 1e-5
    assert relerr(attn.double(), a_ref) < 1e-4 and bool((attn[mask] == 0).all())
    assert relerr(wh[:, :H].double(), w_ref) < 1e-4
    # ---- backward ----
    dwh = r(B, 2 * H)
    d_h1x, d_c1 = r(B, H), r(B, H)
    d_gates, d_c0, dlogit = torch.empty(B, 4 * H, device=dev), torch.empty(B, H, device=dev), torch.empty(B, L, device=dev)
    ops._call("vln_envdrop_ctx_step_bwd", ops._ptr(ctx), ops._ptr(cw), ops._ptr(lengths), ops._ptr(attn), ops._ptr(dwh), 2 * H,
              ops._ptr(dlogit), ops._ptr(acts), ops._ptr(c0), ops._ptr(c1), ops._ptr(d_h1x), ops._ptr(d_c1), ops._ptr(d_gates),
              ops._ptr(d_c0), B, L, H, p, rng.ptr, off, ops._stream())
    a64 = attn.double()
    rr = torch.einsum("blh,bh->bl", ctx.double(), dwh[:, :H].double())
    dl_ref = a64 * (rr - (a64 * rr).sum(1, keepdim=True))
    dl_ref = dl_ref.masked_fill(mask, 0.0)
    dhd = dwh[:, H:].double() + torch.einsum("bl,blh->bh", dl_ref, cw.double())
    dh = dhd * keep + d_h1x.double()
    tc = torch.tanh(c1.double())
    ia, fa, ga, oa = acts.double().chunk(4, 1)
    dc = d_c1.double() + dh * oa * (1 - tc * tc)
    dg_ref = torch.cat((dc * ga * ia * (1 - ia), dc * c0.double() * fa * (1 - fa), dc * ia * (1 - ga * ga), dh * tc * oa * (1 - oa)), 1)
    assert relerr(dlogit.double(), dl_ref) < 1e-4
    assert relerr(d_gates.double(), dg_ref) < 1e-4
    assert relerr(d_c0.double(), dc * fa) < 1e-4


@pytest.mark.parametrize("p", [0.0, 0.5])
@pytest.mark.parametrize("M", [64, 128, 37])
def test_linear_state_epilogues_equal_separate_kernels(setup, M, p):
    """vln_linear_state_fwd / _bwd (GEMM + tile epilogue, csrc/gemm.cu) == vln_linear_bf16x3 followed by
    vln_envdrop_state_fwd / _bwd, bit for bit (same reductions into y are not ordered, so y itself to 1e-6)."""
    import ctypes as C
    world, store, ops, dev = setup
    H, F, KX, OH = 512, 2176, 2752, 2240
    g = torch.Generator().manual_seed(M)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    rng = ops.Rng(5, dev)
    ctr = torch.zeros(16, dtype=torch.int32, device=dev)
    off_q, off_c = 3, 9
    # ---- forward: pre = W_out wh; h~ = tanh(pre) -> xh_next / hq_next / hc ----
    w = r(H, 2 * H) * 0.05
    sw = ops._SplitWeight(w).fresh(w)
    wh = r(M, 2 * H)
    outs = []
    for fused in (False, True):
        pre = torch.zeros(M, H, device=dev)
        xh, hq, hc = torch.zeros(M, KX, device=dev), torch.zeros(M, H, device=dev), torch.zeros(M, H, device=dev)
        if fused:
            for _ in range(2):          # twice: the counters must return to zero
                pre.zero_()
                ops._call("vln_linear_state_fwd", ops._ptr(sw.hi), ops._ptr(sw.lo), H, 2 * H, ops._ptr(wh), 2 * H, M, ops._ptr(pre), H,
                          C.c_void_p(xh.data_ptr() + 4 * OH), KX, ops._ptr(hq), ops._ptr(hc), p, rng.ptr, off_q, off_c, ops._ptr(ctr),
                          ops._stream())
        else:
            ops._call("vln_linear_bf16x3", ops._ptr(sw.hi), ops._ptr(sw.lo), H, 2 * H, ops._ptr(wh), 2 * H, M, None, ops._ptr(pre), H, 1, 0,
                      ops._stream())
            ops._call("vln_envdrop_state_fwd", ops._ptr(pre), 1, C.c_void_p(xh.data_ptr() + 4 * OH), KX, ops._ptr(hq), ops._ptr(hc), M, H, p,
                      rng.ptr, off_q, off_c, ops._stream())
        outs.append((pre, xh, hq, hc))
    assert int(ctr.abs().sum()) == 0
    for a, b in zip(*outs):
        assert relerr(b, a) < 1e-5
    assert torch.equal(outs[0][2] == 0, outs[1][2] == 0) and torch.equal(outs[0][3] == 0, outs[1][3] == 0)     # same masks
    # ---- backward: d_hq = dq W_vin; d_pre = (drop_c'(d_hc) + d_xh_next + drop_q'(d_hq)) * (1 - h~^2) ----
    wv = r(F, H) * 0.05
    sv = ops._SplitWeight(wv).fresh(wv)
    dq, d_hc = r(M, F), r(M, H)
    dxh, xh_t = r(M, KX), torch.tanh(r(M, KX))
    outs = []
    for fused in (False, True):
        dhq, dpre = torch.zeros(M, H, device=dev), torch.zeros(M, H, device=dev)
        if fused:
            ops._call("vln_linear_state_bwd", ops._ptr(sv.hi_t), ops._ptr(sv.lo_t), H, F, ops._ptr(dq), F, M, ops._ptr(dhq), H,
                      ops._ptr(d_hc), C.c_void_p(dxh.data_ptr() + 4 * OH), KX, C.c_void_p(xh_t.data_ptr() + 4 * OH), KX, ops._ptr(dpre),
                      p, rng.ptr, off_q, off_c, ops._ptr(ctr), ops._stream())
        else:
            ops._call("vln_linear_bf16x3", ops._ptr(sv.hi_t), ops._ptr(sv.lo_t), H, F, ops._ptr(dq), F, M, None, ops._ptr(dhq), H, 1, 0,
                      ops._stream())
            ops._call("vln_envdrop_state_bwd", ops._ptr(d_hc), C.c_void_p(dxh.data_ptr() + 4 * OH), KX, ops._ptr(dhq),
                      C.c_void_p(xh_t.data_ptr() + 4 * OH), KX, 1, ops._ptr(dpre), M, H, p, rng.ptr, off_q, off_c, ops._stream())
        outs.append((dhq, dpre))
    assert int(ctr.abs().sum()) == 0
    for a, b in zip(*outs):
        assert relerr(b, a) < 1e-5


@pytest.mark.parametrize("R,M,N", [(2688, 2048, 2752), (5376, 2176, 512), (2688, 512, 1024), (10240, 512, 512), (300, 64, 128),
                                   (77, 132, 260), (2688, 2048, 512)])
def test_wgrad_tcgen05(setup, R, M, N):
    """vln_wgrad_tf32 (csrc/wgrad.cu: dW = dY^T X on tcgen05 kind::tf32, operands MN-major straight from TMA) against
    the float64 product: TF32 inputs, fp32 accumulation -> max-rel <= 2e-3, cosine > 1 - 1e-6; strided operand views,
    accumulation into an existing buffer, ragged edges; identical bits on every run (ordered merge of the row ranges)."""
    world, store, ops, dev = setup
    g = torch.Generator().manual_seed(R + M)
    big_dy = torch.randn(R, M + 8, generator=g).to(dev)
    big_x = torch.randn(R, N + 12, generator=g).to(dev)
    dy, x = big_dy[:, 4:4 + M], big_x[:, 8:8 + N]              # row-strided views with 16-byte aligned rows
    ref = dy.double().t() @ x.double()
    out = ops.wgrad_tc(dy, x)
    assert relerr(out.double(), ref) < 2e-3
    cos = torch.dot(out.double().flatten(), ref.flatten()) / (out.double().norm() * ref.norm())
    assert float(cos) > 1 - 1e-6
    again = ops.wgrad_tc(dy, x)
    assert torch.equal(out, again)
    base = torch.randn(M, N, generator=g).to(dev)
    acc = base.clone()
    ops.wgrad_tc(dy, x, out=acc)
    assert relerr(acc.double(), ref + base.double()) < 2e-3


@pytest.mark.parametrize("M,R,N", [(5120, 1024, 256), (1280, 1024, 256), (300, 132, 260), (129, 64, 128)])
def test_dgrad_tcgen05(setup, M, R, N):
    """vln_dgrad_tf32 (dX = dY W, dY as the K-major tf32 operand of the same tcgen05 kernel) against the float64 product."""
    world, store, ops, dev = setup
    g = torch.Generator().manual_seed(M + N)
    dy = torch.randn(M, R + 4, generator=g).to(dev)[:, :R]
    w = torch.randn(R, N + 8, generator=g).to(dev)[:, 4:4 + N]
    ref = dy.double() @ w.double()
    out = ops.dgrad_tc(dy, w)
    assert relerr(out.double(), ref) < 2e-3
    cos = torch.dot(out.double().flatten(), ref.flatten()) / (out.double().norm() * ref.norm())
    assert float(cos) > 1 - 1e-6
    assert torch.equal(out, ops.dgrad_tc(dy, w))


@pytest.mark.parametrize("n,B,L,H", [(35, 128, 80, 512), (70, 64, 80, 512), (5, 3, 17, 100), (1, 2, 41, 129)])
def test_seq_outer_sum(setup, n, B, L, H):
    """vln_seq_outer_sum (d_ctx over the steps of a rollout) equals the batched fp32 product; strided views; accumulation."""
    world, store, ops, dev = setup
    g = torch.Generator().manual_seed(n * B)
    a = torch.randn(n, B, L, generator=g).to(dev)
    big = torch.randn(n, B, 2 * H, generator=g).to(dev)
    v = big[:, :, H:]
    ref = torch.bmm(a.double().permute(1, 2, 0), v.double().transpose(0, 1))
    out = ops.seq_outer_sum(a, v)
    assert relerr(out.double(), ref) < 1e-5
    base = torch.randn(B, L, H, generator=g).to(dev)
    acc = base.clone()
    ops.seq_outer_sum(a, v, out=acc)
    assert relerr(acc.double(), ref + base.double()) < 1e-5
