"""Container-only: oracle/port_speaker.py against the UNMODIFIED reference Speaker (src/agent/speaker.py, src/model/units.py).

  (a) SpeakerEncoder / SpeakerDecoder modules: eval, train with torch-RNG-matched dropout, gradients;
  (b) the real Speaker class on the reference's own R2RBatch (FakeSim world): from_shortest_path, teacher_forcing
      (train loss + gradients, eval metrics, for_listener scores through the features= entry point of beam search),
      greedy infer_batch.  The class reads EnvDrop-original observation keys its repository's env does not produce, so
      the env is wrapped by an adapter that only renames keys (OldKeys below); .cuda() is the identity here (CPU box).
Also writes nothing: the golden vectors for the GPU tests come from oracle/make_golden.py."""
import random
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import clvln_b200  # noqa: E402,F401
from clvln_b200.environ import make_items, make_world  # noqa: E402
from oracle import port_env as PE, port_modules as P, port_speaker as PS, ref_harness as H, ref_loader  # noqa: E402


def close(a, b, name, tol=1e-5):
    d = (a - b).abs().max().item()
    rel = d / max(b.abs().max().item(), 1e-12)
    print(f"  {name}: abs {d:.3e} rel {rel:.3e}")
    assert rel < tol, name


def grads(params):
    return torch.cat([p.grad.flatten() if p.grad is not None else torch.zeros_like(p).flatten() for p in params])


def leaf_sd(module):
    return {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in module.state_dict().items()}


# ---------------------------------------------------------------- (a) modules
U, _ = ref_loader.load_ref_models()
torch.manual_seed(0)
B, T, Lw, V = 4, 5, 9, 60
for bi in (True, False):
    enc = U.SpeakerEncoder(2176, 512, 0.6, bi, 128, 0.3)
    dec = U.SpeakerDecoder(V, 256, 0, 512, 0.6)
    esd, dsd = leaf_sd(enc), leaf_sd(dec)
    can = torch.randn(B, T, 2176) * 0.3
    img = torch.randn(B, T, 36, 2176) * 0.3
    words = torch.randint(4, V, (B, Lw))
    words[:, 0] = 3
    words[1, 6:] = 0
    lengths = [5, 3, 4, 1]
    cmask = PS.length_mask(lengths, T)
    for mode in ("eval", "train"):
        enc.train(mode == "train"), dec.train(mode == "train")
        enc.zero_grad(), dec.zero_grad()
        torch.manual_seed(11)
        ctx = enc(can.clone(), img.clone(), lengths)
        z = torch.zeros(1, B, 512)
        logit, h1, c1 = dec(words, ctx, cmask, z, z)
        (logit.square().mean() + h1.sum() * 1e-3).backward()
        g_ref = grads(list(enc.parameters()) + list(dec.parameters()))
        torch.manual_seed(11)
        drop = P.Drop("torch") if mode == "train" else None
        for v in list(esd.values()) + list(dsd.values()):
            v.grad = None
        ctx2 = PS.speaker_encoder(esd, can, img, bidirectional=bi, drop=drop)
        logit2, h12, c12 = PS.speaker_decoder(dsd, words, ctx2, cmask, z, z, drop=drop)
        (logit2.square().mean() + h12.sum() * 1e-3).backward()
        names = [n for n, _ in enc.named_parameters()] + [n for n, _ in dec.named_parameters()]
        sds = [esd] * len(list(enc.parameters())) + [dsd] * len(list(dec.parameters()))
        g_port = grads([sd[n] for sd, n in zip(sds, names)])
        print(f"modules bi={bi} {mode}")
        close(ctx2, ctx, "encoder ctx")
        close(logit2, logit, "decoder logit", 1e-4)
        close(h12, h1, "h1")
        close(c12, c1, "c1")
        close(g_port, g_ref, "gradients", 1e-4)
    # step-wise decoding with a carried state == the full sequence
    enc.eval(), dec.eval()
    ctx = enc(can.clone(), img.clone(), lengths)
    full, _, _ = PS.speaker_decoder(dsd, words, ctx, cmask, z, z)
    h, c = z, z
    for t in range(Lw):
        step, h, c = PS.speaker_decoder(dsd, words[:, t:t + 1], ctx, cmask, h, c)
        close(step[:, 0], full[:, t].detach(), f"step {t}", 1e-4)

# ---------------------------------------------------------------- (b) the Speaker class on the reference env
w = make_world(n_scans=3, seed=1)
items = make_items(w, 40, seed=1)
src = H.install(w, {"train": items})
import src.agent.speaker as ref_speaker  # noqa: E402
import src.environ as environ  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self            # CPU container: the class hard-codes .cuda()
ref_speaker.Variable = torch.autograd.Variable            # (speaker.py:422 uses Variable without importing it)
if not hasattr(np, "bool"):
    np.bool = bool                                        # (speaker.py:335 uses the alias numpy 1.24 removed)


class OldKeys:
    """The reference's own R2RBatch under the observation keys its Speaker class reads (speaker.py:125-226)."""

    def __init__(self, r2r):
        self.r2r, self.env = r2r, r2r.env
        self.feature_size, self.batch_size = r2r.feature_size, r2r.batch_size

    def reset(self, **kw):
        return self._wrap(self.r2r.reset(**kw))

    def _get_obs(self):
        return self._wrap(self.r2r.observe())

    def reset_epoch(self, **kw):
        return self.r2r.reset_epoch(**kw)

    def size(self):
        return self.r2r.size()

    @staticmethod
    def _wrap(obs):
        out = []
        for ob in obs:
            o = dict(ob)
            o["viewpoint"] = ob["viewpointId"]
            o["candidate"] = [dict(c, pointId=c["absViewIndex"], viewpointId=c["nextViewpointId"]) for c in ob["candidates"]]
            out.append(o)
        return out


tok = H.StubTokenizer(items)
fs = H.feature_store(w)
view = PE.WorldView(w)
spk_cfg = ref_loader.AttrDict(RNN_DIM=512, DROPOUT=0.6, FEAT_DROPOUT=0.3, BI_DIRECTION=True, WEMB=256, LR=1e-4,
                              FAST_TRAIN=False, IGNORE_ID=-1, MAX_DECODE=24, LOAD_OPTIM=False)
random.seed(2020)
torch.manual_seed(2020)
renv = environ.R2RBatch(fs, batch_size=8, splits=["train"], tokenizer=tok)
H.warm_candidate_buffer(renv)
spk = ref_speaker.Speaker(spk_cfg, torch.device("cpu"), tok, env=OldKeys(renv))
random.seed(2020)
penv = PE.R2RBatchPort(view, items, batch_size=8)
assert [d["instr_id"] for d in penv.data] == [d["instr_id"] for d in renv.data]
port = PS.SpeakerPort(leaf_sd(spk.encoder), leaf_sd(spk.decoder), max_decode=24)
for it in range(2):
    robs = spk.env.reset()
    pobs = penv.reset()
    # -- path features
    (img_r, can_r), len_r = spk.from_shortest_path()
    (img_p, can_p), len_p, _ = PS.from_shortest_path(penv, pobs)
    print(f"batch {it}: path lengths {len_r.tolist()}")
    assert (np.asarray(len_r) == len_p).all()
    close(img_p, img_r, "img_feats", 1e-7)
    close(can_p, can_r, "can_feats", 1e-7)
    insts = torch.from_numpy(np.array([ob["instr_encoding"] for ob in robs]))
    feats = ((img_r, can_r), len_r)
    # -- teacher forcing: train loss + gradients (torch-RNG-matched dropout), through the env and through features=
    spk.env.reset(restart=True)
    spk.encoder.zero_grad(), spk.decoder.zero_grad()
    torch.manual_seed(5)
    loss_r = spk.teacher_forcing(train=True)
    loss_r.backward()
    g_ref = grads(list(spk.encoder.parameters()) + list(spk.decoder.parameters()))
    for v in list(port.enc.values()) + list(port.dec.values()):
        v.grad = None
    torch.manual_seed(5)
    loss_p = port.teacher_forcing(((img_p, can_p), len_p), insts, train=True, drop=P.Drop("torch"))
    loss_p.backward()
    names = [("enc", n) for n, _ in spk.encoder.named_parameters()] + [("dec", n) for n, _ in spk.decoder.named_parameters()]
    g_port = grads([getattr(port, k)[n] for k, n in names])
    close(loss_p.detach(), loss_r.detach(), "teacher-forcing loss (train)")
    close(g_port, g_ref, "gradients", 1e-4)
    # -- eval metrics and the per-word scores beam search asks for
    spk.env.reset(restart=True)
    l_r, wa_r, sa_r = spk.teacher_forcing(train=False)
    l_p, wa_p, sa_p, _ = port.teacher_forcing(((img_p, can_p), len_p), insts, train=False)
    print(f"  eval loss {l_r:.6f} / {l_p:.6f}  word acc {wa_r:.4f} / {wa_p:.4f}  sent acc {sa_r} / {sa_p}")
    assert abs(l_r - l_p) < 1e-5 * max(1.0, abs(l_r)) and abs(wa_r - wa_p) < 1e-9 and sa_r == sa_p
    spk.encoder.eval(), spk.decoder.eval()
    sc_r = spk.teacher_forcing(train=False, features=feats, insts=insts, for_listener=True)
    sc_p = port.teacher_forcing(((img_p, can_p), len_p), insts, train=False, for_listener=True)
    close(sc_p, sc_r.detach(), "for_listener scores", 1e-4)
    # -- greedy decoding
    spk.env.reset(restart=True)
    words_r = spk.infer_batch()
    words_p, _ = port.infer_batch(((img_p, can_p), len_p))
    print("  greedy words", words_r.shape, words_r[0][:8].tolist())
    assert words_r.shape == words_p.shape and (words_r == words_p).all()
print("speaker port OK")
