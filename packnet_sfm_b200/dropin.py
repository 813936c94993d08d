"""Register the kernel-backed classes under the reference's own module paths.

The reference resolves its components by module path + class name (packnet_sfm/utils/load.py:79-105,
model_wrapper.py:382-408) and hard-imports the loss (SelfSupModel.py:4).  install() therefore substitutes

    packnet_sfm.networks.depth.PackNet01                 -> packnet_sfm_b200.networks  (class PackNet01)
    packnet_sfm.networks.layers.packnet.layers01         -> packnet_sfm_b200.networks  (layer classes)
    packnet_sfm.losses.multiview_photometric_loss        -> packnet_sfm_b200.losses    (class MultiViewPhotometricLoss)

in sys.modules BEFORE the reference's models are imported, so SfmModel / SelfSupModel / ModelWrapper /
scripts/train.py run unchanged on the sm_100a kernels.  The rest of `packnet_sfm` (models, geometry.Pose,
datasets, trainers, config) stays the reference's own code."""
import importlib
import importlib.machinery
import sys
import types


def install(reference_root=None):
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from . import losses, networks

    def alias(name, source, attrs):
        mod = types.ModuleType(name)
        mod.__pn_b200_dropin__ = True
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)   # load_class() probes importlib.util.find_spec
        mod.__doc__ = "packnet_sfm_b200 drop-in for %s" % name
        for a in attrs:
            setattr(mod, a, getattr(source, a))
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        try:
            setattr(importlib.import_module(parent), leaf, mod)
        except Exception:  # parent package not importable yet: sys.modules entry is enough
            pass
        return mod

    alias("packnet_sfm.networks.layers.packnet.layers01", networks,
          ["Conv2D", "ResidualConv", "ResidualBlock", "InvDepth", "PackLayerConv3d", "UnpackLayerConv3d"])
    alias("packnet_sfm.networks.depth.PackNet01", networks, ["PackNet01"])
    alias("packnet_sfm.losses.multiview_photometric_loss", losses, ["MultiViewPhotometricLoss"])
    return True
