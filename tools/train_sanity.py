"""End-to-end sanity of the whole optimisation loop on the GPU: 300 graph-replayed EnvDrop iterations (paired rollouts,
A2C + imitation, clip + RMSprop) on a 6-scan synthetic world; the teacher-forced imitation loss in eval mode goes down.
usage: python tools/train_sanity.py"""
import random, sys, torch
sys.path.insert(0, "/root/repo")
import clvln_b200
from clvln_b200 import utils
from clvln_b200.agent import build_agent
from clvln_b200.engine import build_trainer
from clvln_b200.environ import make_world, make_items, R2RBatch
dev = torch.device("cuda:0")
world = make_world(n_scans=6, seed=3, device=dev)
items = make_items(world, 512, seed=3)
cfg = utils.agent_cfg("ENVDROP")
cfg.TRAIN.BATCH_SIZE = 64
cfg.TRAIN.LR = 1e-4
random.seed(2020)
env = R2RBatch(world, items, batch_size=64, device=dev)
torch.manual_seed(2020)
agent = build_agent(cfg, utils.StubTokenizer(), dev)
agent.env = env
agent.train()
step = build_trainer(cfg, env, dev).make_step(cfg, agent)
print(type(step).__name__)
ml = []
for it in range(300):
    loss = step()
    if it % 50 == 49:
        agent.eval()
        env.reset_epoch(shuffle=False)
        with torch.no_grad():
            agent.rollout(train_ml=True, train_rl=False, feedback="teacher")
        agent.train()
        print(it + 1, "iteration loss %.4f" % float(loss), "| teacher-forced imitation loss (eval) %.4f" % float(agent.loss["ml_loss"]))
