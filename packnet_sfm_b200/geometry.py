"""The one piece of pose algebra the hot path needs on the host: PoseNet's 6-vector -> [B,4,4] rigid transform.

The fused loss consumes `Pose.mat` and returns dL/dmat, so this module only has to (a) hold the matrix under the
attribute name the reference's loss reads (`pose.mat`, packnet_sfm/geometry/pose.py:18) and (b) build it from
[tx,ty,tz,rx,ry,rz] with the reference's convention R = Rx(rx)·Ry(ry)·Rz(rz) (pose_utils.py:8-37), differentiably.
The rotation is written in closed form (one stack of nine products of sines / cosines) instead of three batched
matrix products.  With the reference on the path (dropin.install()) its own `packnet_sfm.geometry.pose.Pose` is used
and this shim is not imported by the glue."""
import torch


def rigid_from_vec(vec):
    """[B,6] (translation, XYZ Euler angles) -> [B,4,4]."""
    t, a = vec[:, :3], vec[:, 3:]
    s, c = torch.sin(a), torch.cos(a)
    sx, sy, sz = s.unbind(1)
    cx, cy, cz = c.unbind(1)
    sxsy, cxsy = sx * sy, cx * sy
    zero, one = torch.zeros_like(sx), torch.ones_like(sx)
    rows = [cy * cz, -(cy * sz), sy, t[:, 0],
            sxsy * cz + cx * sz, cx * cz - sxsy * sz, -(sx * cy), t[:, 1],
            sx * sz - cxsy * cz, cxsy * sz + sx * cz, cx * cy, t[:, 2],
            zero, zero, zero, one]
    return torch.stack(rows, 1).view(-1, 4, 4)


class Pose:
    """Holder of a [B,4,4] transform under the reference's attribute name (`.mat`)."""

    def __init__(self, mat):
        if mat.dim() == 2:
            mat = mat.unsqueeze(0)
        if mat.dim() != 3 or tuple(mat.shape[-2:]) != (4, 4):
            raise ValueError("Pose needs a [B,4,4] tensor, got %s" % (tuple(mat.shape),))
        self.mat = mat

    def __len__(self):
        return len(self.mat)

    @classmethod
    def identity(cls, N=1, device=None, dtype=torch.float):
        return cls(torch.eye(4, device=device, dtype=dtype).expand(N, 4, 4).contiguous())

    @classmethod
    def from_vec(cls, vec, mode="euler"):
        if mode != "euler":
            raise ValueError("Rotation mode not supported {}".format(mode))
        return cls(rigid_from_vec(vec))

    def to(self, *args, **kwargs):
        self.mat = self.mat.to(*args, **kwargs)
        return self
