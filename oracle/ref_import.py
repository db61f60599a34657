"""TEST INFRASTRUCTURE ONLY -- imports the *real* PhysDock reference from /root/reference.

Only `oracle/make_golden.py` and the `not gpu` pin tests (skipped when the reference tree is
absent, as it is on the GPU box) may use this module.  Nothing in the product package
`physdock_b200/` imports it.

The reference needs two packages this image lacks (`ml_collections`, `rdkit`); both are shimmed with
in-process stand-ins that the hot path never exercises:
  * `ml_collections.ConfigDict`  -> attribute-access dict (used by PhysDock/configs.py:195)
  * `rdkit.*`                    -> empty modules (imported at PhysDock/models/model.py:19-21; only
                                    touched by the MMFF branch model.py:26-52, which we never run)
The slow scipy truncated-normal initialiser (PhysDock/models/primitives/linear.py:33-44) is replaced
with a plain normal init: weights are always overwritten by `physdock_b200.synthetic.make_dit_state`.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PHYSDOCK_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "PhysDock", "models"))


class _ConfigDict(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = _ConfigDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


_done = False


def install_shims():
    global _done
    if _done:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "ml_collections" not in sys.modules:
        m = types.ModuleType("ml_collections")
        m.ConfigDict = _ConfigDict
        sys.modules["ml_collections"] = m
    for n in ["rdkit", "rdkit.Chem", "rdkit.Chem.AllChem", "rdkit.Geometry", "rdkit.rdBase",
              "rdkit.Chem.rdchem", "rdkit.Chem.rdmolops"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.modules["rdkit.Chem"].AllChem = sys.modules["rdkit.Chem.AllChem"]
    sys.modules["rdkit.Geometry"].Point3D = object
    sys.modules["rdkit.rdBase"].DisableLog = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # fast init: patch before the layer modules bind the names
    import torch
    import PhysDock.models.primitives.linear as lin  # noqa

    def _fast_trunc(weights, scale=1.0, fan="fan_in"):
        f = lin._calculate_fan(weights.shape, fan)
        with torch.no_grad():
            weights.normal_(0.0, (scale / max(1, f)) ** 0.5)

    lin.trunc_normal_init_ = _fast_trunc
    _done = True


def import_reference():
    """Returns the reference's (PhysDock, PhysDockConfig, AF3DiT, tensor_utils module)."""
    install_shims()
    import torch
    prec = torch.get_float32_matmul_precision()
    from PhysDock.models.model import PhysDock  # sets matmul precision "high" (model.py:5)
    from PhysDock.configs import PhysDockConfig
    from PhysDock.models.layers.transformers import AF3DiT
    import PhysDock.utils.tensor_utils as tu
    torch.set_float32_matmul_precision(prec)  # CPU matmuls ignore it; keep global state clean
    return PhysDock, PhysDockConfig, AF3DiT, tu


def build_reference_dit(model_name: str = "medium"):
    """The reference denoiser `AF3DiT` (PhysDock/models/layers/transformers.py:178) with the
    dimension table of PhysDock/configs.py:59-88 for `model_name`."""
    _, PhysDockConfig, AF3DiT, _ = import_reference()
    cfg = PhysDockConfig(model_name=model_name)
    return AF3DiT(**cfg.model.dit).float().eval()
