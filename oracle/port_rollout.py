"""Oracle port of the three agents' rollout()/loss bodies — TEST INFRASTRUCTURE ONLY.

Restates envdrop.py:86-278, follower.py:65-173 and monitor.py:89-199 on top of
oracle/port_modules.py (weights = the reference's state_dicts) and any environment with the
reference's obs-dict protocol (oracle/port_env.py, or the real R2RBatch via ref_harness).
Semantics kept: CUDA-device behaviour of `.cpu()` (copy), float64 A2C
intermediates (envdrop.py:209-212, 243-252), `ended.all()` early exit, per-agent loss
reductions.  Dead reference paths (speaker, avoid_cyclic) are not restated.

`feedback` may also be a list/array of forced actions per step (``forced[t][i]``), which is
how a CUDA rollout's sampled actions are replayed through the oracle.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import port_modules as P


class Agent:
    """Bundle of state_dicts + hyper-parameters (configs/*/*.yaml MODEL.<NAME>)."""

    def __init__(self, kind, encoder_sd, decoder_sd, critic_sd=None, *, hidden, bidirectional,
                 enc_layers, drop_rate=0.5, feat_drop_rate=0.3, episode_len=10, ml_weight=0.2,
                 gamma=0.9, rl_normalize="total", device="cpu"):
        self.kind = kind
        self.enc, self.dec, self.cri = encoder_sd, decoder_sd, critic_sd
        self.hidden, self.bi, self.layers = hidden, bidirectional, enc_layers
        self.p, self.pf = drop_rate, feat_drop_rate
        self.episode_len, self.ml_weight, self.gamma = episode_len, ml_weight, gamma
        self.rl_normalize = rl_normalize
        self.device = device
        self.training = False
        self.logs = {"entropy": [], "critic_loss": [], "total": []}
        self.trace = {}

    def params(self):
        out = []
        for sd in (self.enc, self.dec, self.cri):
            if sd is not None:
                out += [v for v in sd.values() if v.requires_grad]
        return out


# ---- observation marshalling (base.py:114-178, envdrop.py:75-84, monitor.py:68-87) ----
def instr_tensors(obs, full_length, device):
    seq = np.array([ob["instr_encoding"] for ob in obs])
    lengths = torch.from_numpy(np.array([ob["instr_length"] for ob in obs]))
    seq = torch.from_numpy(seq)
    if not full_length:
        seq = seq[:, :int(lengths[0])]
    return seq.long().to(device), (seq == 0).to(device), lengths


def pano_tensor(obs, device):
    return torch.from_numpy(np.stack([ob["feature"] for ob in obs]).astype(np.float32)).to(device)


def cand_tensor(obs, device, fdim=2176):
    lens = [len(ob["candidates"]) + 1 for ob in obs]
    out = np.zeros((len(obs), max(lens), fdim), np.float32)
    for i, ob in enumerate(obs):
        for j, c in enumerate(ob["candidates"]):
            out[i, j] = c["feature"]
    return torch.from_numpy(out).to(device), lens


def pose_tensor(obs, device):
    from .port_env import angle_feat
    return torch.from_numpy(np.stack([angle_feat(ob["heading"], ob["elevation"]) for ob in obs])).to(device)


def teacher_actions(obs, ended):
    a = np.zeros(len(obs), np.int64)
    for i, ob in enumerate(obs):
        if ended[i]:
            a[i] = -1
            continue
        a[i] = len(ob["candidates"])
        for k, c in enumerate(ob["candidates"]):
            if c["nextViewpointId"] == ob["teacher"]:
                a[i] = k
                break
    return a


def length_mask(lens, device):                                      # misc.py:481-486
    lens = torch.as_tensor(lens)
    return (torch.arange(int(lens.max())).unsqueeze(0) >= lens.unsqueeze(1)).to(device)


def _choose(feedback, t, logits, target):
    extra = {}
    if isinstance(feedback, str):
        if feedback == "teacher":
            return target, extra
        if feedback == "argmax":
            return logits.max(1)[1], extra
        if feedback == "sample":
            dist = torch.distributions.Categorical(F.softmax(logits, 1))
            return dist.sample(), extra
        raise NotImplementedError(feedback)
    return torch.as_tensor(np.asarray(feedback[t]), device=logits.device).long(), extra


def _env_actions(a_t, obs, ended):
    cpu = a_t.detach().cpu().numpy().copy()
    for i, a in enumerate(cpu):
        if a == len(obs[i]["candidates"]) or a == -1 or ended[i]:
            cpu[i] = -1
    return cpu


def _traj0(obs):
    return [{"instr_id": ob["instr_id"], "path": [(ob["viewpointId"], ob["heading"], ob["elevation"])]}
            for ob in obs]


# ------------------------------------------------------------------------------------
def rollout_envdrop(ag, env, *, train_ml=True, train_rl=False, train_cl=False, restart=False,
                    feedback="sample", drop=None, record=None):
    """envdrop.py:86-278.  Returns (traj, loss dict)."""
    dev = ag.device
    is_sample = (feedback == "sample") or (not isinstance(feedback, str) and train_rl)
    if isinstance(feedback, str) and feedback != "sample":
        train_rl = False
    obs = env.reset(restart=restart)
    B = len(obs)
    seq, seq_mask, lengths = instr_tensors(obs, False, dev)
    ctx, h_t, c_t = P.encoder_lstm(ag.enc, seq, lengths, bidirectional=ag.bi, num_layers=ag.layers,
                                   drop_ratio=ag.p, drop=drop)
    traj = _traj0(obs)
    ended = np.zeros(B, bool)
    last_dist = np.array([ob["distance"] for ob in obs], np.float32)
    rewards, hiddens, logps, masks, ents = [], [], [], [], []
    ml = torch.zeros(B, device=dev) if train_cl else 0.0
    rl = torch.zeros(B, device=dev) if train_cl else 0.0
    h_tilde = h_t
    steps = []

    def decode(obs, h_tilde, c_t):
        pose = pose_tensor(obs, dev)
        img = pano_tensor(obs, dev)
        cand, lens = cand_tensor(obs, dev)
        logit, (h1, c1), ht, _ = P.envdrop_decoder(
            ag.dec, pose, img, cand, h_tilde, c_t, ctx, seq_mask,
            drop_ratio=ag.p, feat_drop_ratio=ag.pf, drop=drop)
        return logit, h1, c1, ht, lens

    for t in range(ag.episode_len):
        logit, h_t, c_t, h_tilde, lens = decode(obs, h_tilde, c_t)
        hiddens.append(h_t)
        logit = logit.masked_fill(length_mask(lens, dev), -float("inf"))
        target = torch.from_numpy(teacher_actions(obs, ended)).to(dev)
        ce = F.cross_entropy(logit, target, ignore_index=-1, reduction="none")
        ml = ml + (ce if train_cl else ce.sum())
        if is_sample:                                               # envdrop.py:188-194
            cat = torch.distributions.Categorical(F.softmax(logit, 1))
            a_t = cat.sample() if isinstance(feedback, str) else \
                torch.as_tensor(np.asarray(feedback[t]), device=dev).long()
            logps.append(cat.log_prob(a_t.clamp(min=0)))
            ent = cat.entropy()
            ents.append(ent)
            ag.logs["entropy"].append(ent.sum().item())
        else:
            a_t, _ = _choose(feedback, t, logit, target)
            if not (isinstance(feedback, str) and feedback == "teacher"):   # envdrop.py:184-187
                logps.append(F.log_softmax(logit, 1).gather(1, a_t.clamp(min=0).unsqueeze(1)))
        steps.append(dict(logits=logit.detach().clone(), target=target.clone(), action=a_t.detach().clone()))
        cpu_a = _env_actions(a_t, obs, ended)
        obs = env.step(cpu_a, obs, traj)
        dist = np.array([ob["distance"] for ob in obs], np.float32)
        stop = cpu_a == -1
        reward = (stop * (2 * (dist < 3) - 1) * 2 + (1 - stop) * np.sign(last_dist - dist)) * (~ended)
        rewards.append(reward)
        masks.append(~ended)
        last_dist[:] = dist
        ended[:] = np.logical_or(ended, stop)
        if ended.all():
            break

    if train_rl:
        _, last_h, _, _, _ = decode(obs, h_tilde, c_t)
        with torch.no_grad():
            last_v = P.critic(ag.cri, last_h, ag.p, drop).detach().cpu().numpy()
        disc = (~ended) * last_v
        total = 0
        for t in range(len(rewards) - 1, -1, -1):
            disc = disc * ag.gamma + rewards[t]
            m = torch.from_numpy(masks[t]).to(dev)
            r = torch.from_numpy(disc).to(dev)
            v = P.critic(ag.cri, hiddens[t], ag.p, drop)
            adv = (r - v).detach()
            cur = torch.zeros(B, device=dev)
            cur += (-logps[t] * adv * m)
            cur += (((r - v) ** 2) * m) * 0.5
            if is_sample:
                cur += (-0.01 * ents[t] * m)
            rl = rl + (cur if train_cl else cur.sum())
            ag.logs["critic_loss"].append((((r - v) ** 2) * m).sum().item())
            total = total + np.sum(masks[t])
        ag.logs["total"].append(total)
        if ag.rl_normalize == "total":
            rl = rl / total
        elif ag.rl_normalize == "batch":
            rl = rl / B
    ag.trace = dict(steps=steps, rewards=rewards, masks=masks)
    loss = {"ml_loss": ml * ag.ml_weight / B if train_ml else 0.0, "rl_loss": rl if train_rl else 0.0}
    return traj, loss


def rollout_follower(ag, env, *, train_cl=False, restart=False, feedback="sample", drop=None):
    """follower.py:65-173.  Returns (traj, ml_loss)."""
    dev = ag.device
    obs = env.reset(restart=restart)
    B = len(obs)
    seq, seq_mask, lengths = instr_tensors(obs, False, dev)
    ctx, h_t, c_t = P.encoder_lstm(ag.enc, seq, lengths, bidirectional=ag.bi, num_layers=ag.layers,
                                   drop_ratio=ag.p, drop=drop)
    traj = _traj0(obs)
    a_prev = torch.zeros(B, 2176, device=dev)
    ended = np.zeros(B, bool)
    ml = torch.zeros(B, device=dev) if train_cl else 0.0
    steps = []
    for t in range(ag.episode_len):
        img = pano_tensor(obs, dev)
        cands, lens = cand_tensor(obs, dev)
        logit, (h_t, c_t), _ = P.follower_decoder(ag.dec, img, a_prev, cands, h_t, c_t, ctx, seq_mask,
                                                  drop_ratio=ag.p, drop=drop)
        logit = logit.masked_fill(length_mask(lens, dev), -float("inf"))
        target = torch.from_numpy(teacher_actions(obs, ended)).to(dev)
        if train_cl:
            ml = ml + F.cross_entropy(logit, target, ignore_index=-1, reduction="none")
        else:
            ml = ml + F.cross_entropy(logit, target, ignore_index=-1)
        a_t, _ = _choose(feedback, t, logit, target)
        steps.append(dict(logits=logit.detach().clone(), target=target.clone(), action=a_t.detach().clone()))
        cpu_a = _env_actions(a_t, obs, ended)
        obs = env.step(cpu_a, obs, traj)
        stop = cpu_a == -1
        a_prev = cands[np.arange(B), np.maximum(cpu_a, 0), :].detach()
        ended[:] = np.logical_or(ended, stop)
        if ended.all():
            break
    ag.trace = dict(steps=steps)
    return traj, ml


def rollout_monitor(ag, env, *, train_cl=False, restart=False, feedback="sample", lamb=0.5, drop=None):
    """monitor.py:89-199.  Returns (traj, ml_loss, progress_loss)."""
    dev = ag.device
    obs = env.reset(restart=restart)
    B = len(obs)
    seq, seq_mask, lengths = instr_tensors(obs, True, dev)
    ctx, h_t, c_t = P.encoder_lstm(ag.enc, seq, lengths, bidirectional=ag.bi, num_layers=ag.layers,
                                   drop_ratio=ag.p, drop=drop)
    traj = _traj0(obs)
    a_prev = torch.zeros(B, 2176, device=dev)
    ended = np.zeros(B, bool)
    start_dist = np.array([ob["distance"] for ob in obs], np.float32)
    cur_dist = start_dist.copy()
    ml, prog_log = 0.0, 0.0
    steps = []
    for t in range(ag.episode_len):
        cands, lens = cand_tensor(obs, dev)
        cmask = length_mask(lens, dev)
        (logit, prog), (h_t, c_t), _ = P.monitor_decoder(
            ag.dec, a_prev, cands, h_t, c_t, ctx, seq_mask, cmask, drop_ratio=ag.p,
            training=ag.training, drop=drop)
        logit = logit.masked_fill(cmask, -float("inf"))
        target = torch.from_numpy(teacher_actions(obs, ended)).to(dev)
        act_loss = F.cross_entropy(logit, target, ignore_index=-1,
                                   reduction="none" if train_cl else "mean")
        if t == 0:
            cur = act_loss
        else:
            pt = (start_dist - cur_dist) / start_dist
            pt[cur_dist <= 3.0] = 1.0
            pt[ended] = prog.detach().cpu().numpy()[ended]
            pt = torch.from_numpy(pt).to(dev)
            pl = F.mse_loss(prog, pt, reduction="none" if train_cl else "mean")
            prog_log += pl.mean().item()
            cur = lamb * pl + (1 - lamb) * act_loss
        ml = ml + cur
        a_t, _ = _choose(feedback, t, logit, target)
        steps.append(dict(logits=logit.detach().clone(), target=target.clone(), action=a_t.detach().clone(),
                          progress=prog.detach().clone()))
        cpu_a = _env_actions(a_t, obs, ended)
        obs = env.step(cpu_a, obs, traj)
        stop = cpu_a == -1
        cur_dist[:] = np.array([ob["distance"] for ob in obs], np.float32)
        ended[:] = np.logical_or(ended, stop)
        a_prev = cands[np.arange(B), np.maximum(cpu_a, 0), :].detach()
        if ended.all():
            break
    ag.trace = dict(steps=steps)
    return traj, ml, prog_log
