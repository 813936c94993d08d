"""Build container only: the drop-in registration makes the UNMODIFIED reference glue resolve to our classes."""
import sys

import pytest

pytestmark = pytest.mark.ref


def test_reference_models_pick_up_the_dropin_classes():
    from oracle import ref_shims
    ref_shims.install()
    for name in list(sys.modules):
        if name.startswith("packnet_sfm.models") or name in ("packnet_sfm.losses.multiview_photometric_loss",
                                                             "packnet_sfm.networks.depth.PackNet01"):
            del sys.modules[name]
    from packnet_sfm_b200 import dropin, losses, networks
    dropin.install()
    try:
        from packnet_sfm.models.SelfSupModel import SelfSupModel        # reference file, unmodified
        from packnet_sfm.utils.load import load_class                   # reference plug-in mechanism
        model = SelfSupModel(num_scales=4, photometric_reduce_op="min", automask_loss=True, clip_loss=0.0)
        assert type(model._photometric_loss) is losses.MultiViewPhotometricLoss
        cls = load_class("PackNet01", paths=["packnet_sfm.networks.depth"])
        assert cls is networks.PackNet01
        net = cls(version="1A", dropout=0.0)
        assert len(net.state_dict()) == 216
    finally:
        ref_shims.install()     # drops the aliases again so later tests see the real reference
        for name in list(sys.modules):
            if name.startswith("packnet_sfm.models"):
                del sys.modules[name]
