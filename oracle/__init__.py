"""Oracle — TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy / plain PyTorch fp32) of the reference's rollout hot path
(IMNearth/Curriculum-Learning-For-VLN, tasks/R2R-judy/src).  Nothing in the
product package imports this directory; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may.

Parity status: the reference ships NO golden vectors for this path
 — "parity unpinned" by the reference's own tests.  The port is
instead pinned against outputs of the reference itself, run in the build
container by oracle/make_golden.py (imports /root/reference through
oracle/ref_loader.py) and committed as fixtures under tests/golden/.
"""
