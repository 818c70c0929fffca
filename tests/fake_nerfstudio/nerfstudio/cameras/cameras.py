import torch


class Cameras:
    """Batch of pinhole cameras, shape [N] (or () after integer indexing, like nerfstudio's TensorDataclass)."""

    def __init__(self, camera_to_worlds, fx, fy, cx, cy, width, height, _shape=None):
        c2w = torch.as_tensor(camera_to_worlds, dtype=torch.float32)
        self._shape = tuple(c2w.shape[:-2]) if _shape is None else tuple(_shape)
        n = 1
        for s in self._shape:
            n *= s
        self.camera_to_worlds = c2w.reshape(self._shape + (3, 4))

        def t(v, dtype):
            v = torch.as_tensor(v, dtype=dtype)
            return v.expand(self._shape + (1,)).clone() if v.dim() == 0 else v.reshape(self._shape + (1,)).clone()

        self.fx, self.fy, self.cx, self.cy = (t(v, torch.float32) for v in (fx, fy, cx, cy))
        self.width, self.height = t(width, torch.int64), t(height, torch.int64)
        self.metadata = None
        self.distortion_params = None

    @property
    def shape(self):
        return self._shape

    def __len__(self):
        if not self._shape:
            raise TypeError("len() of a 0-d Cameras")
        return self._shape[0]

    def _sub(self, f):
        c = Cameras.__new__(Cameras)
        c.camera_to_worlds = f(self.camera_to_worlds)
        for k in ("fx", "fy", "cx", "cy", "width", "height"):
            setattr(c, k, f(getattr(self, k)))
        c._shape = tuple(c.camera_to_worlds.shape[:-2])
        c.metadata, c.distortion_params = None, None
        return c

    def __getitem__(self, i):
        assert self._shape, "cannot index a 0-d Cameras"
        return self._sub(lambda x: x[i])        # int -> shape (), slice -> shape [n]

    def reshape(self, shape):
        shape = tuple(shape)
        return self._sub(lambda x: x.reshape(shape + tuple(x.shape[len(self._shape):])))

    def to(self, device):
        return self._sub(lambda x: x.to(device))

    def rescale_output_resolution(self, scaling_factor):
        s = float(scaling_factor)
        self.fx, self.fy, self.cx, self.cy = self.fx * s, self.fy * s, self.cx * s, self.cy * s
        self.height = (self.height * s).to(torch.int64)
        self.width = (self.width * s).to(torch.int64)
