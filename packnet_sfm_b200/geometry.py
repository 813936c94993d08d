"""Host-side mirror of the reference's tiny pose helpers (PyTorch, differentiable).

  Pose                <- packnet_sfm/geometry/pose.py:9-101
  euler2mat           <- packnet_sfm/geometry/pose_utils.py:8-37   (R = Rx @ Ry @ Rz)
  pose_vec2mat        <- packnet_sfm/geometry/pose_utils.py:41-51
  invert_pose         <- packnet_sfm/geometry/pose_utils.py:55-60

north_star keeps PoseNet and the pose algebra in PyTorch host code; the fused loss kernel consumes
Pose.mat ([B,4,4]) directly and returns its gradient, so autograd carries on into these ops."""
import torch


def euler2mat(angle):
    B = angle.size(0)
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    cosz, sinz = torch.cos(z), torch.sin(z)
    zeros = z.detach() * 0
    ones = zeros.detach() + 1
    zmat = torch.stack([cosz, -sinz, zeros, sinz, cosz, zeros, zeros, zeros, ones], dim=1).view(B, 3, 3)
    cosy, siny = torch.cos(y), torch.sin(y)
    ymat = torch.stack([cosy, zeros, siny, zeros, ones, zeros, -siny, zeros, cosy], dim=1).view(B, 3, 3)
    cosx, sinx = torch.cos(x), torch.sin(x)
    xmat = torch.stack([ones, zeros, zeros, zeros, cosx, -sinx, zeros, sinx, cosx], dim=1).view(B, 3, 3)
    return xmat.bmm(ymat).bmm(zmat)


def pose_vec2mat(vec, mode="euler"):
    if mode is None:
        return vec
    trans, rot = vec[:, :3].unsqueeze(-1), vec[:, 3:]
    if mode != "euler":
        raise ValueError("Rotation mode not supported {}".format(mode))
    return torch.cat([euler2mat(rot), trans], dim=2)


def invert_pose(T):
    Tinv = torch.eye(4, device=T.device, dtype=T.dtype).repeat([len(T), 1, 1])
    Tinv[:, :3, :3] = torch.transpose(T[:, :3, :3], -2, -1)
    Tinv[:, :3, -1] = torch.bmm(-1.0 * Tinv[:, :3, :3], T[:, :3, -1].unsqueeze(-1)).squeeze(-1)
    return Tinv


class Pose:
    """[B,4,4] rigid transform wrapper with the reference's interface."""

    def __init__(self, mat):
        assert tuple(mat.shape[-2:]) == (4, 4)
        if mat.dim() == 2:
            mat = mat.unsqueeze(0)
        assert mat.dim() == 3
        self.mat = mat

    def __len__(self):
        return len(self.mat)

    @classmethod
    def identity(cls, N=1, device=None, dtype=torch.float):
        return cls(torch.eye(4, device=device, dtype=dtype).repeat([N, 1, 1]))

    @classmethod
    def from_vec(cls, vec, mode):
        mat = pose_vec2mat(vec, mode)
        pose = torch.eye(4, device=vec.device, dtype=vec.dtype).repeat([len(vec), 1, 1])
        pose[:, :3, :3] = mat[:, :3, :3]
        pose[:, :3, -1] = mat[:, :3, -1]
        return cls(pose)

    @property
    def shape(self):
        return self.mat.shape

    def item(self):
        return self.mat

    def repeat(self, *args, **kwargs):
        self.mat = self.mat.repeat(*args, **kwargs)
        return self

    def inverse(self):
        return Pose(invert_pose(self.mat))

    def to(self, *args, **kwargs):
        self.mat = self.mat.to(*args, **kwargs)
        return self

    def transform_pose(self, pose):
        assert tuple(pose.shape[-2:]) == (4, 4)
        return Pose(self.mat.bmm(pose.item()))

    def transform_points(self, points):
        assert points.shape[1] == 3
        B, _, H, W = points.shape
        out = self.mat[:, :3, :3].bmm(points.view(B, 3, -1)) + self.mat[:, :3, -1].unsqueeze(-1)
        return out.view(B, 3, H, W)

    def __matmul__(self, other):
        if isinstance(other, Pose):
            return self.transform_pose(other)
        if isinstance(other, torch.Tensor):
            if other.shape[1] == 3 and other.dim() > 2:
                return self.transform_points(other)
            raise ValueError("Unknown tensor dimensions {}".format(other.shape))
        raise NotImplementedError()
