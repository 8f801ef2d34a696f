"""Training loops for TRAIN.CLMODE = "" / "NAIVE" / "SELF-PACE".

Interface of the reference's engine (src/engine/__init__.py:6-17, trainer.py:46,200,361,
curriculum.py:27-179,183-456): ``trainer.train(cfg, agent, tsboard_dir, train_env, valid_env)``
with ``ClassicTrainer``, ``NaiveCurriculum(switch_epoch)`` and ``SelfPacedCurriculum(train_env,
device, pace_func, init_lamb, init_weight_ctrl, miu, interval, strategy, burn_in)``.

What changed underneath: the per-iteration body is a ``TrainStep`` — rollouts on the device,
one backward, one flat-buffer all-reduce + fused clip/update (engine/optim.py) — and nothing in
it reads a value back to the host; losses are kept as device scalars and fetched once per epoch.
The self-paced weights live on the device and are updated by the same tensor program as
curriculum.py:428-448, bit-reproducible given identical losses.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .optim import build_optimizer


def _is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class TrainStep:
    """One optimisation step of an agent (the loop bodies of trainer.py:102-113, 265-270,
    411-429 and curriculum.py:78-102, 283-314)."""

    def __init__(self, cfg, agent, optimizer=None, weights=None):
        self.cfg, self.agent = cfg, agent
        self.opt = optimizer if optimizer is not None else build_optimizer(cfg, agent)
        self.name = cfg.MODEL.NAME
        self.feedback = cfg.AGENT.FEEDBACK
        self.weights = weights              # SelfPacedCurriculum instance (train_cl) or None
        self.item_loss = None
        # EnvDrop: step the teacher-forced and the sampled rollout of an iteration as one batch (agent.rollout_pair)
        self.pair_rollouts = os.environ.get("VLN_PAIR_ROLLOUTS", "1") != "0"
        # the fused rollout's weight-gradient GEMMs run on a side stream under the encoder's backward (agent/fused.py)
        if getattr(agent, "_fused", None) is not None and agent.device.type == "cuda":
            agent._fused.async_wgrad = os.environ.get("VLN_ASYNC_WGRAD", "1") != "0"
        if agent.device.type == "cuda":           # same for the encoder's leaf gradients (ops._leaf_grads_async)
            from .. import ops
            ops.ASYNC_LEAF_GRADS[0] = os.environ.get("VLN_ASYNC_WGRAD", "1") != "0"
        # data parallel: the decoder's + critic's gradient bucket is all-reduced as soon as the decoder's backward through
        # time has produced it, under the encoder's backward (engine/optim.py reduce_range); groups = [encoder, decoder, critic]
        fd = getattr(agent, "_fused", None)
        if (fd is not None and getattr(self.opt, "world", 1) > 1 and fd.async_wgrad and self.name == "ENVDROP"
                and os.environ.get("VLN_EARLY_ALLREDUCE", "1") != "0"):
            lo, hi = self.opt.group_offsets[1], self.opt.grad.numel()
            fd.on_grads_ready = lambda: self.opt.reduce_range(lo, hi)

    def losses(self):
        """Run the rollouts; return (loss to differentiate, per-item loss record or None)."""
        ag, cl = self.agent, self.weights is not None
        if self.name == "ENVDROP" and self.feedback == "sample" and self.pair_rollouts and getattr(ag, "fused", False) \
                and ag.device.type == "cuda" and ag.trace is None:
            ag.rollout_pair(train_cl=cl)                       # both rollouts of the iteration as one batch of 2B episodes
            cur = ag.loss["ml_loss"] + ag.loss["rl_loss"]
            if not cl:
                return cur, None
            w = self.weights.weight[ag.last_batch.index]
            return torch.dot(w, cur), ag.loss["ml_loss"].detach() * cur.shape[0]
        if self.name == "ENVDROP":
            ag.rollout(train_ml=True, train_rl=False, train_cl=cl, feedback="teacher")
            ml = ag.loss["ml_loss"]
            rl = 0.0
            if self.feedback == "sample":
                ag.rollout(train_ml=False, train_rl=True, train_cl=cl, restart=True, feedback="sample")
                rl = ag.loss["rl_loss"]
            cur = ml + rl
            if not cl:
                return cur, None
            w = self.weights.weight[ag.last_batch.index]
            return torch.dot(w, cur), ml.detach() * cur.shape[0]              # curriculum.py:294-296, 312
        if self.name == "SELF-MONITOR" and not cl:
            ag.rollout(train_ml=True, feedback=self.feedback, lamb=self.cfg.TRAIN.PROGMONITOR_WEIGHT)
        else:
            ag.rollout(train_ml=True, train_cl=cl, feedback=self.feedback)
        cur = ag.ml_loss
        if not cl:
            return cur, None
        w = self.weights.weight[ag.last_batch.index]
        return torch.dot(w, cur) / w.sum(), cur.detach()                       # curriculum.py:298-301, 314

    def __call__(self):
        ag = self.agent
        nvtx = torch.cuda.nvtx if ag.device.type == "cuda" else None      # ranges for nsys / ncu --nvtx
        if ag.rng is not None:
            ag.rng.begin_iteration()
        self.opt.zero_grad()
        if nvtx:
            nvtx.range_push("vln/rollouts")
        loss, item = self.losses()
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("vln/backward")
        loss.backward()
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("vln/allreduce+clip+update")
        self.opt.step()
        if nvtx:
            nvtx.range_pop()
        if item is not None:
            self.weights.record(ag.last_batch.index, item)
        return loss.detach()


class ClassicTrainer:
    """TRAIN.CLMODE == "" (src/engine/__init__.py:6-17 dispatching to trainer.py)."""

    def __init__(self, hooks=None):
        self.hooks = hooks or {}

    def pick_env(self, train_env, ep):
        return train_env

    def make_step(self, cfg, agent, weights=None):
        """The iteration body: CUDA-graph replay on a GPU (engine/graphs.py) — the fused EnvDrop rollout, and the module
        path of Follower / Self-Monitor (their rollouts are free of host read-backs: device-side candidate width, fixed
        step count with masks) — the eager TrainStep otherwise (VLN_TRAIN_GRAPH=0, CPU-side tests)."""
        graphable = getattr(agent, "fused", False) if cfg.MODEL.NAME == "ENVDROP" else \
            os.environ.get("VLN_TRAIN_GRAPH_MODULES", "1") != "0"
        if graphable and agent.device.type == "cuda" and os.environ.get("VLN_TRAIN_GRAPH", "1") != "0":
            from .graphs import GraphedTrainStep
            agent.sync_every = 0                  # fixed-length sampled rollouts (ended episodes are masked): no host polls
            return GraphedTrainStep(cfg, agent, weights=weights)
        return TrainStep(cfg, agent, weights=weights)

    def after_epoch(self, ep, step):
        pass

    def train(self, cfg, agent, tsboard_dir, train_env, valid_env, evaluator=None, log=print):
        """``trainer.train(cfg, agent, cfg.OUTPUT.TSBOARD_DIR, train_env, valid_env)`` exactly as main.py:125 calls it:
        with validation envs and no ``evaluator`` the reference's own scoring (Evaluation(...).score of
        trainer.py:69-76, 481-511; engine/evaluator.py here) runs every EVAL_INTERVAL epochs and the best / latest
        checkpoints are written under OUTPUT.CKPT_DIR."""
        tc = cfg.TRAIN
        if valid_env and evaluator is None:
            from .evaluator import evaluate as _evaluate
            if agent.device.type == "cuda":           # the whole split in one kernel launch on the HBM distance table
                evaluator = lambda env, results: _evaluate(env, results, store=agent.store_of(env))
            else:
                evaluator = _evaluate
        rank0 = not _is_dist() or dist.get_rank() == 0
        start = tc.START_EPOCH
        if cfg.OUTPUT.RESUME:
            ckpt = agent.load_model(os.path.join(cfg.OUTPUT.CKPT_DIR, f"{cfg.OUTPUT.RESUME}.pt"))
            if "last_epoch" in ckpt:
                # Monitor/EnvDrop resume at last_epoch+1 (trainer.py:217,378); the others at last_epoch (:63)
                bump = isinstance(self, ClassicTrainer) and type(self) is ClassicTrainer and \
                    cfg.MODEL.NAME in ("SELF-MONITOR", "ENVDROP")
                start = ckpt["last_epoch"] + (1 if bump else 0)
        step = None                                   # built at the first iteration (it allocates the flat GPU buffers)
        best = {k: 0.0 for k in (valid_env or {})}
        history = []
        t0 = time.time()
        for ep in range(start, tc.MAX_EPOCH + 1):
            agent.env = self.pick_env(train_env, ep)
            agent.train()
            agent.reset_loss()
            rec = []
            for it in range(tc.ITER_PER_EPOCH):
                if step is None:
                    step = self.make_step(cfg, agent)
                if hasattr(step, "prefetch_next"):           # stage the next minibatch under this iteration's GPU work,
                    step.prefetch_next = it + 1 < tc.ITER_PER_EPOCH      # except across the epoch boundary (eval / env switch)
                rec.append(step())
            rec = torch.stack(rec).cpu().numpy()                     # the epoch's only loss read-back
            info = {"epoch": ep, "loss_sum": float(rec.sum()), "loss_avg": float(rec.mean()),
                    "loss_min": float(rec.min()), "loss_max": float(rec.max()),
                    "minutes": (time.time() - t0) / 60.0}
            if cfg.MODEL.NAME == "ENVDROP" and agent.logs.get("total"):
                total = max(float(torch.stack(agent.logs["total"]).sum()), 1.0)
                info["critic_loss"] = float(torch.stack(agent.logs["critic_loss"]).sum()) / total
                info["policy_entropy"] = float(torch.stack(agent.logs["entropy"]).sum()) / total
                info["total_actions"] = total
            if valid_env and evaluator is not None and ep % tc.EVAL_INTERVAL == 0:
                agent.eval()
                for key, env in valid_env.items():
                    agent.env = env
                    # data parallel: every rank scores the WHOLE split (an unsharded view of its env), so all ranks
                    # consume the shared `random` stream identically, agree on `best`, and only rank 0 writes files
                    with _unsharded(env):
                        agent.test(iters=None, feedback="argmax")
                    scores = evaluator(env, agent.get_results())
                    info[key] = scores
                    if scores["success_rate"] > best.get(key, 0.0) and cfg.OUTPUT.CKPT_DIR:
                        best[key] = scores["success_rate"]
                        if rank0:
                            _clean_dir(cfg.OUTPUT.CKPT_DIR, f"best_{key}")
                            agent.save_model(os.path.join(cfg.OUTPUT.CKPT_DIR,
                                                          "best_{}_SR:{:.4f}.pt".format(key, scores["success_rate"])),
                                             cfg=cfg, last_epoch=ep)
            self.after_epoch(ep, step)
            if cfg.OUTPUT.CKPT_DIR and rank0:
                _clean_dir(cfg.OUTPUT.CKPT_DIR, "latest_avgloss")
                agent.save_model(os.path.join(cfg.OUTPUT.CKPT_DIR, "latest_avgloss:{:.4f}.pt".format(info["loss_avg"])),
                                 cfg=cfg, last_epoch=ep)
            if cfg.OUTPUT.CKPT_DIR and _is_dist():
                dist.barrier()                      # nobody races ahead of rank 0's checkpoint files
            history.append(info)
            log(f"\t Epoch [{ep}/{tc.MAX_EPOCH}] loss sum {info['loss_sum']:.4f} avg {info['loss_avg']:.4f} "
                f"min {info['loss_min']:.4f} max {info['loss_max']:.4f} ({info['minutes']:.2f} min)")
        self.history = history
        return agent


class NaiveCurriculum(ClassicTrainer):
    """TRAIN.CLMODE == "NAIVE" (curriculum.py:27-179): train_env is {round_k: env over rounds 1..k};
    the round switches every ``switch_epoch`` epochs (:176-179)."""

    def __init__(self, switch_epoch=20, reverse=False, hooks=None):
        super().__init__(hooks)
        self.switch_epoch, self.reverse = switch_epoch, reverse

    def curriculum_strategy(self, train_env, cur_epoch):
        idx = 1 + (cur_epoch - 1) // self.switch_epoch
        return train_env[f"round_{idx}"] if idx <= 4 else train_env["round_5"]

    def pick_env(self, train_env, ep):
        return self.curriculum_strategy(train_env, ep)


class SelfPacedCurriculum(ClassicTrainer):
    """TRAIN.CLMODE == "SELF-PACE" (Jiang et al., AAAI 2015; curriculum.py:183-456)."""

    def __init__(self, train_env, device, pace_func="linear", init_lamb=0.1, init_weight_ctrl=0.5, miu=0.1,
                 interval=5, strategy="epoch", burn_in=10, hooks=None):
        super().__init__(hooks)
        self.train_env, self.device, self.pace_func = train_env, device, pace_func
        self.dim = len(train_env)
        self.a = torch.from_numpy(train_env.a).to(device)
        self.c = torch.tensor(train_env.c, device=device)
        self.lamb = torch.tensor(init_lamb, device=device)
        self.weight = self._init_weight_(init_weight_ctrl)
        self.stepsize, self.burn_in = miu, burn_in
        self.update_interval, self.update_strategy = interval, strategy
        self.loss_for_item = torch.zeros(self.dim, device=device)

    def _init_weight_(self, val):                                   # curriculum.py:214-220
        w = torch.ones(self.dim, device=self.device) * val
        w[self.a <= 2] = 1.0
        return w

    def pick_env(self, train_env, ep):
        return self.train_env

    def make_step(self, cfg, agent, weights=None):
        return super().make_step(cfg, agent, weights=self)

    def record(self, index, item_loss):
        """loss_for_item[idx] = loss (curriculum.py:311-314).  Under data parallelism every rank
        gathers all ranks' (idx, loss) pairs so the weight vector stays replicated."""
        if _is_dist():
            ws = dist.get_world_size()
            idx_all = [torch.empty_like(index) for _ in range(ws)]
            loss_all = [torch.empty_like(item_loss) for _ in range(ws)]
            dist.all_gather(idx_all, index)
            dist.all_gather(loss_all, item_loss)
            index, item_loss = torch.cat(idx_all), torch.cat(loss_all)
        self.loss_for_item[index] = item_loss

    def after_epoch(self, ep, step):                                # curriculum.py:402-416
        if ep >= self.burn_in and self.update_interval and ep % self.update_interval == 0:
            if bool(self.lamb < self.loss_for_item.max()):
                self.lamb = self.lamb + self.stepsize
            else:
                self.lamb = self.lamb + self.stepsize / 2
            self.update_weight(self.loss_for_item)

    def update_weight(self, loss):
        if self.update_strategy != "epoch":
            raise NotImplementedError
        self._update_epoch_(loss)

    def _update_epoch_(self, epoch_loss):                           # curriculum.py:428-448
        zeta = 1 - self.lamb
        mask = epoch_loss >= self.lamb
        w = self.weight
        w[mask] = 0.01
        if self.pace_func == "log":
            w[~mask] = torch.log(epoch_loss[~mask] + zeta) / torch.log(zeta)
        elif self.pace_func == "linear":
            w[~mask] = 1 - epoch_loss[~mask] / self.lamb
        elif self.pace_func == "binary":
            w[~mask] = 1.0
        else:
            raise NotImplementedError
        w[w < 0.01] = 0.01
        if torch.dot(self.a, w) > self.c:
            a_norm = torch.norm(self.a, p=2)
            new_w = w + self.a * (self.c - torch.dot(self.a, w)) / (a_norm * a_norm)
            new_w[new_w <= 0.0] = 0.001
            self.weight.copy_(new_w)        # in place: a captured CUDA graph keeps reading this very tensor


def build_trainer(cfg, train_env, device):
    """main.py:96-124: pick the trainer from TRAIN.CLMODE."""
    mode = cfg.TRAIN.CLMODE
    if mode == "NAIVE":
        return NaiveCurriculum()
    if mode == "SELF-PACE":
        sp = cfg.TRAIN.SELF_PACE
        return SelfPacedCurriculum(train_env, device, pace_func=sp.FUNC, init_lamb=sp.LAMB, init_weight_ctrl=sp.WCTRL,
                                   miu=sp.MIU, interval=sp.INTERVAL, strategy=sp.STRATEGY, burn_in=sp.BURN_IN)
    return ClassicTrainer()


class _unsharded:
    """Temporarily make a data-parallel env hand out whole (world_size = 1) minibatches."""

    def __init__(self, env):
        self.env = env

    def __enter__(self):
        e = self.env
        self.saved = (getattr(e, "rank", 0), getattr(e, "world_size", 1))
        if self.saved[1] > 1:
            e.rank, e.world_size = 0, 1
        return e

    def __exit__(self, *exc):
        if self.saved[1] > 1:
            self.env.rank, self.env.world_size = self.saved
        return False


def _clean_dir(save_dir, key):
    if not os.path.isdir(save_dir):
        os.makedirs(save_dir, exist_ok=True)
        return
    for fn in os.listdir(save_dir):
        if key in fn:
            os.remove(os.path.join(save_dir, fn))
