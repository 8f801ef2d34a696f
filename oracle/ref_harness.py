"""Container-only harness: run the UNMODIFIED reference env + agents on a synthetic World.

TEST INFRASTRUCTURE.  The reference's R2RBatch / CLR2RBatch / EnvBatch / make_candidate /
agents (common_env.py, curriculum_env.py, src/agent/*.py) execute as they are; only their
*external inputs* are substituted:

  MatterSim.Simulator      -> FakeSim below: a discretised graph walker over World tables
  utils.load_nav_graphs    -> networkx graphs built from World.edge_len  (no connectivity/*.json)
  utils.load_datasets      -> the synthetic episode items                 (no R2R json)
  Tokenizer                -> StubTokenizer: returns the items' pre-made token ids

FakeSim semantics: viewIndex in 0..35, heading=(v%12)*pi/6, elevation=(v//12-1)*pi/6; a
neighbour is navigable only from the one view World assigns it (rel_heading = World.cand_rel,
rel_elevation = 0), listed after the current location in ascending neighbour order —
so make_candidate's 36-view sweep yields World's candidate order, absViewIndex,
normalized_heading (= state.heading + rel_heading, the same float64 sum World stores)
and loc_elevation bit for bit.
"""
import types

import networkx as nx
import numpy as np

from . import ref_loader


def _wmod():
    import clvln_b200
    return clvln_b200.environ.world


class _Loc:
    __slots__ = ("viewpointId", "rel_heading", "rel_elevation", "ix")

    def __init__(self, vid, rh=0.0, re=0.0, ix=0):
        self.viewpointId, self.rel_heading, self.rel_elevation, self.ix = vid, rh, re, ix


class _State:
    pass


def make_fake_sim_class(world):
    W = _wmod()
    scan_index = {s: i for i, s in enumerate(world.scans)}
    vp_index = [{n: j for j, n in enumerate(names)} for names in world.vp_names]

    class FakeSim:
        def __init__(self):
            self.g = None
            self.view = 12

        # configuration calls made by M3DSimulator.new (misc.py:346-362)
        def setRenderingEnabled(self, *_): pass
        def setDiscretizedViewingAngles(self, *_): pass
        def setCameraResolution(self, *_): pass
        def setCameraVFOV(self, *_): pass
        def init(self): pass

        def newEpisode(self, scanId, viewpointId, heading, elevation):
            s = scan_index[scanId]
            self.s = s
            self.g = world.gid(s, vp_index[s][viewpointId])
            row = 1 + int(round(elevation / W.ANGLE_INC))
            self.view = row * 12 + int(round(heading / W.ANGLE_INC)) % 12

        def _navigable(self):
            g = self.g
            s = self.s
            locs = [_Loc(world.vp_names[s][int(world.vp_local[g])])]
            for j in range(int(world.n_cand[g])):
                if int(world.cand_view[g, j]) == self.view:
                    v = int(world.cand_vp[g, j])
                    locs.append(_Loc(world.vp_names[s][int(world.vp_local[v])],
                                     float(world.cand_rel[g, j]), 0.0, v))
            return locs

        def getState(self):
            st = _State()
            st.scanId = world.scans[self.s]
            st.location = _Loc(world.vp_names[self.s][int(world.vp_local[self.g])])
            st.viewIndex = self.view
            st.heading = W.view_heading(self.view)
            st.elevation = W.view_elevation(self.view)
            st.navigableLocations = self._navigable()
            return st

        def makeAction(self, index, heading, elevation):
            if index:
                self.g = self._navigable()[index].ix
                return
            h = (self.view % 12 + int(heading)) % 12
            r = min(2, max(0, self.view // 12 + int(elevation)))
            self.view = r * 12 + h

    return FakeSim


class StubTokenizer:
    """Returns the synthetic items' token ids for their (unique) instruction strings."""

    def __init__(self, items, vocab=992):
        self.table = {it["instructions"]: (it["instr_encoding"], it["instr_length"]) for it in items}
        self._vocab = vocab
        self.word_to_index = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<BOS>": 3}

    def vocab_size(self):
        return self._vocab

    def encode_sentence(self, sentence, **_):
        enc, n = self.table[sentence]
        return np.array(enc), n

    def split_sentence(self, s):
        return s.split()


def ref_items(items):
    """Items in the reference's json shape: one entry per path, `instructions` a list."""
    out = []
    for it in items:
        d = {k: it[k] for k in ("scan", "path_id", "path", "heading", "distance")}
        d["instructions"] = [it["instructions"]]
        out.append(d)
    return out


def install(world, datasets):
    """Patch the reference's external inputs.  ``datasets``: {split_name: [items]} keyed the
    way load_datasets is called ("train", "val_seen", "train_round[1]_v3", ...)."""
    src = ref_loader.load_ref_agents()
    import src.utils.misc as misc
    import src.utils as utils
    import MatterSim
    MatterSim.Simulator = make_fake_sim_class(world)

    def load_nav_graphs(scans):
        graphs = {}
        for scan in scans:
            s = world.scans.index(scan)
            G = nx.Graph()
            for (u, v), w in world.edge_len[s].items():
                G.add_edge(world.vp_names[s][u], world.vp_names[s][v], weight=w)
            graphs[scan] = G
        return graphs

    def load_datasets(splits, dataset="R2R", data_dir=None):
        data = []
        for sp in splits:
            data += ref_items(datasets[sp])
        return data

    for m in (misc, utils):
        m.load_nav_graphs = load_nav_graphs
        m.load_datasets = load_datasets
    return src


def feature_store(world):
    """The reference's in-RAM feature dict: "scan_vp" -> np.float32[36, 2048] (misc.py:254-279),
    holding the bf16-rounded table values upcast to fp32."""
    t = world.table.float().cpu().numpy()
    return {world.long_id(g): t[g] for g in range(world.n_vp)}


def warm_candidate_buffer(env):
    """make_candidate's first visit to a viewpoint associates its heading arithmetic differently
    from every later (buffered) visit (common_env.py:249-256 vs :283-285).  Visit every
    viewpoint once so that all observations come from the steady-state buffered path."""
    feats = env.env.features
    for long_id, f in feats.items():
        scan, vp = long_id.split("_", 1)
        if scan in env.scans:
            env.make_candidate(f, scan, vp, 12)


def model_cfg(name):
    """MODEL.<NAME> nodes of configs/*/*.yaml as attribute dicts."""
    A = ref_loader.AttrDict
    if name == "ENVDROP":
        return A(WORD_EMB_SIZE=256, ACT_EMB_SIZE=64, HIDDEN_SIZE=512, DROP_RATE=0.5,
                 FEAT_DROP_RATE=0.3, ENC_BIDIRECTION=True, ENC_LAYERS=1, ML_WEIGHT=0.2,
                 GAMMA=0.9, RL_NORMALIZE="total")
    if name == "FOLLOWER":
        return A(WORD_EMB_SIZE=300, HIDDEN_SIZE=256, DROP_RATE=0.5, ENC_BIDIRECTION=True,
                 ENC_LAYERS=2)
    if name == "MONITOR":
        return A(WORD_EMB_SIZE=256, HIDDEN_SIZE=512, DROP_RATE=0.5, ENC_BIDIRECTION=False,
                 ENC_LAYERS=1, MLP_HIDDEN=[1024], PROGMONITOR_WEIGHT=0.5)
    raise KeyError(name)
