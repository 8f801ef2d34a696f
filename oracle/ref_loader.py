"""Container-only loader for the *real* reference (read-only at /root/reference).

TEST INFRASTRUCTURE.  Used by oracle/make_golden.py and by the `not gpu` tests
that pin oracle/port_*.py against the reference when /root/reference exists.
It never travels to the GPU box (the reference directory is absent there).

Two levels:
  * load_ref_models()  -> (units, policy) modules; they import only torch.
  * load_ref_agents()  -> the `src` package with dummy sys.modules entries for
    the un-installed imports (MatterSim, yacs, prettytable, tensorboardX, boto3).
"""
import importlib
import importlib.util
import os
import sys
import types

import torch

# the reference itself where it exists (this container), else its unmodified copy under oracle/_ref (oracle/build_ref.py;
# what reaches the GPU box)
REF_ROOT = os.environ.get("VLN_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, "tasks", "R2R-judy", "src", "model", "units.py")):
    REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_TASK = os.path.join(REF_ROOT, "tasks", "R2R-judy")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_TASK, "src", "model", "units.py"))


def load_ref_models():
    """units.py / policy.py under a synthetic package (policy.py:8 does `from . import units`)."""
    if "refmodel.policy" in sys.modules:
        return sys.modules["refmodel.units"], sys.modules["refmodel.policy"]
    pkg = types.ModuleType("refmodel")
    pkg.__path__ = [os.path.join(REF_TASK, "src", "model")]
    sys.modules["refmodel"] = pkg
    mods = []
    for name in ("units", "policy"):
        spec = importlib.util.spec_from_file_location(
            f"refmodel.{name}", os.path.join(REF_TASK, "src", "model", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refmodel.{name}"] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


class _AttrDict(dict):
    """Stand-in for yacs.config.CfgNode: attribute access + clone()."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_cpu_patched = False


def patch_cpu_copy():
    """Reference bug: on a CPU device `a_t.detach().cpu().numpy()`
    aliases the action tensor, and the rollout then writes -1 into tensors saved
    for backward.  On CUDA `.cpu()` copies; reproduce that semantics."""
    global _cpu_patched
    if _cpu_patched:
        return
    orig = torch.Tensor.cpu

    def cpu(t, *a, **k):
        out = orig(t, *a, **k)
        return out.clone() if t.device.type == "cpu" else out
    torch.Tensor.cpu = cpu
    _cpu_patched = True


def load_ref_agents():
    """Import the reference's `src` package (agents, engine) with stubs."""
    if "src.agent" in sys.modules and getattr(sys.modules["src"], "__ref__", False):
        return sys.modules["src"]

    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return lambda *a, **kw: None

    _stub_module("MatterSim", Simulator=_Dummy)
    _stub_module("prettytable", PrettyTable=_Dummy)
    yacs = _stub_module("yacs")
    yc = _stub_module("yacs.config", CfgNode=_AttrDict)
    yacs.config = yc
    _stub_module("boto3")
    be = _stub_module("botocore.exceptions", ClientError=Exception)
    bc = _stub_module("botocore")
    bc.exceptions = be
    _stub_module("tensorboardX", SummaryWriter=_Dummy)
    if REF_TASK not in sys.path:
        sys.path.insert(0, REF_TASK)
    patch_cpu_copy()
    src = importlib.import_module("src")
    src.__ref__ = True
    importlib.import_module("src.agent")
    importlib.import_module("src.engine")
    return src


AttrDict = _AttrDict
