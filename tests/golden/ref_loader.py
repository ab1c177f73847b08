"""Import the reference's hot-path modules from /root/reference (THIS container only).

Used by `make_golden.py` and by the optional oracle-vs-reference test; never on the GPU box
(`/root/reference` does not exist there). Only the modules SURVEY.md §8(c) lists are imported:
`vgqa.core.decoder.*`, `vgqa.core.language.bert_module`, `vgqa.core.model_utils`,
`vgqa.core.vision.position_encoding`, `vgqa.core.postprocessor`, `vgqa.training.evaluator`
(functions only), `vgqa.utils.{training_utils,box_ops}`.

The package `__init__` files of `vgqa`, `vgqa.core`, `vgqa.core.language`, `vgqa.core.vision`,
`vgqa.training`, `vgqa.utils` pull in decord / ffmpeg / timm / transformers / torchtext, none of which
the hot path needs, so those packages are registered as bare namespace modules (their `__init__` is
not executed) and `easydict` gets a 6-line stand-in.
"""
import importlib
import os
import sys
import types
from types import SimpleNamespace

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_COMPILED = os.path.join(_REPO, "oracle", "_ref")     # tools/make_oracle_ref.py: the same modules byte-compiled (travels to the GPU box)


def _pick_root() -> str:
    env = os.environ.get("VGQA_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/vgqa/core/decoder"):
        return "/root/reference"
    return _COMPILED


REF_ROOT = _pick_root()


def reference_available() -> bool:
    """The reference SOURCE tree is mounted (build container): golden makers / live comparisons may run."""
    return os.path.isdir(os.path.join(REF_ROOT, "vgqa", "core", "decoder")) and REF_ROOT != _COMPILED


def reference_modules_available() -> bool:
    """The reference's hot-path modules can be imported: from the source tree, or from `oracle/_ref` (byte-compiled by
    tools/make_oracle_ref.py in the build container; that is what exists on the GPU box)."""
    return os.path.isdir(os.path.join(REF_ROOT, "vgqa", "core", "decoder")) and \
        (REF_ROOT != _COMPILED or os.path.isfile(os.path.join(REF_ROOT, "vgqa", "core", "decoder", "__init__.bin")))


def _namespace_pkg(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, mod)
    return mod


class _CompiledFinder:
    """Meta-path finder for `oracle/_ref`: `vgqa.a.b` → oracle/_ref/vgqa/a/b.bin (or .../b/__init__.bin), a marshalled code
    object written by tools/make_oracle_ref.py with the same Python minor version."""

    def __init__(self, root):
        self.root = root
        info = open(os.path.join(root, "BUILD_INFO")).readline().split()
        want = f"{sys.version_info[0]}.{sys.version_info[1]}"
        if len(info) < 2 or info[1] != want:
            raise RuntimeError(f"oracle/_ref was compiled for python {info[1:2]} but this is python {want}: rebuild it")

    def find_spec(self, name, path=None, target=None):
        import importlib.util
        if not name.startswith("vgqa."):
            return None
        base = os.path.join(self.root, *name.split("."))
        for file, is_pkg in ((os.path.join(base, "__init__.bin"), True), (base + ".bin", False)):
            if os.path.isfile(file):
                spec = importlib.util.spec_from_loader(name, self, origin=file, is_package=is_pkg)
                if is_pkg:
                    spec.submodule_search_locations = [base]
                return spec
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        with open(module.__spec__.origin, "rb") as f:
            code = marshal.loads(f.read())
        exec(code, module.__dict__)


def _install_shims():
    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                self.__dict__ = self

        ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except Exception:
            tq = types.ModuleType("tqdm")
            tq.tqdm = lambda x, *a, **k: x
            sys.modules["tqdm"] = tq


def load_reference():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if not reference_modules_available():
        raise RuntimeError(f"reference modules not found at {REF_ROOT} (run tools/make_oracle_ref.py where /root/reference is mounted)")
    _install_shims()
    if REF_ROOT == _COMPILED and not any(isinstance(f, _CompiledFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _CompiledFinder(REF_ROOT))
    v = os.path.join(REF_ROOT, "vgqa")
    _namespace_pkg("vgqa", v)
    _namespace_pkg("vgqa.core", os.path.join(v, "core"))
    _namespace_pkg("vgqa.core.language", os.path.join(v, "core", "language"))
    _namespace_pkg("vgqa.core.vision", os.path.join(v, "core", "vision"))
    _namespace_pkg("vgqa.utils", os.path.join(v, "utils"))
    _namespace_pkg("vgqa.training", os.path.join(v, "training"))

    ns = SimpleNamespace()
    dec = importlib.import_module("vgqa.core.decoder")
    ns.build_encoder = dec.build_encoder
    ns.build_decoder = dec.build_decoder
    ns.build_TemporalSampling = dec.build_TemporalSampling
    ns.build_SpatialActivation = dec.build_SpatialActivation
    mu = importlib.import_module("vgqa.core.model_utils")
    ns.MLP = mu.MLP
    ns.gen_sineembed_for_position = mu.gen_sineembed_for_position
    pe = importlib.import_module("vgqa.core.vision.position_encoding")
    ns.PositionEmbeddingSine = pe.PositionEmbeddingSine
    tu = importlib.import_module("vgqa.utils.training_utils")
    ns.NestedTensor = tu.NestedTensor
    pp = importlib.import_module("vgqa.core.postprocessor")
    ns.PostProcess = pp.PostProcess
    ev = importlib.import_module("vgqa.training.evaluator")
    ns.linear_interp = ev.linear_interp
    ns.linear_interp_conf = ev.linear_interp_conf
    ns.single_forward = ev.single_forward
    return ns


def load_feature_resizer():
    """`FeatureResizer` of the text encoder (vgqa/core/language/bert.py:77-96).  bert.py imports
    `pytorch_pretrained_bert` (absent here, import-time only) — stubbed; `transformers` is installed."""
    load_reference()
    if "pytorch_pretrained_bert" not in sys.modules:
        ppb = types.ModuleType("pytorch_pretrained_bert")
        ppb.__path__ = []
        mod = types.ModuleType("pytorch_pretrained_bert.modeling")
        mod.BertModel = object
        ppb.modeling = mod
        sys.modules["pytorch_pretrained_bert"] = ppb
        sys.modules["pytorch_pretrained_bert.modeling"] = mod
    return importlib.import_module("vgqa.core.language.bert").FeatureResizer


def make_cfg(max_video_len: int = 200, hidden=256, heads=8, ffn=2048, enc_layers=6, dec_layers=6):
    """Attribute-tree stand-in for the yacs cfg: only the keys the hot path reads at construction
    (vgqa/config/defaults.py:7,63-72; SOLVER.USE_ATTN :153)."""
    vstg = SimpleNamespace(HIDDEN=hidden, HEADS=heads, FFN_DIM=ffn, DROPOUT=0.1, ENC_LAYERS=enc_layers,
                           DEC_LAYERS=dec_layers, QUERY_DIM=4, FROM_SCRATCH=True,
                           USE_LEARN_TIME_EMBED=False, USE_ACTION=True)
    return SimpleNamespace(
        INPUT=SimpleNamespace(MAX_VIDEO_LEN=max_video_len),
        MODEL=SimpleNamespace(VSTG=vstg),
        SOLVER=SimpleNamespace(USE_ATTN=False, USE_AUX_LOSS=True),
        DATASET=SimpleNamespace(APP_NUM=20, MOT_NUM=34),
    )
