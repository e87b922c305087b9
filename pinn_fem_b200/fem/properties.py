"""Material properties: constants or small tanh MLPs (reference: fem/properties.py).

``NNProperty`` keeps the reference's constructor and ``value()`` contract, but
evaluates the network with the CUDA kernels (``pf_mlp_forward``) in fp64 -- the
module's parameters are the single source of truth and may live on any device /
dtype; they are flattened to an fp64 device vector for every evaluation."""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np
import torch

from .. import ops


class Property:
    """Interface of a material property: ``value(inputs=None)``."""

    def value(self, inputs: Optional[Any] = None):
        raise NotImplementedError

    def is_trainable(self) -> bool:
        return False

    def get_torch_params(self) -> list:
        return []


class ScalarProperty(Property):
    def __init__(self, value: float):
        self._value = float(value)

    def value(self, inputs: Optional[Any] = None) -> float:
        return self._value

    def __repr__(self) -> str:
        return f"ScalarProperty({self._value:.3e})"


def _linear_stack(net) -> List[torch.nn.Linear]:
    """The Linear layers of a Linear/Tanh stack; raise for anything the kernels do not implement."""
    seq = net.net if hasattr(net, "net") and isinstance(net.net, torch.nn.Sequential) else net
    if not isinstance(seq, torch.nn.Sequential):
        raise NotImplementedError("NNProperty: only torch.nn.Sequential Linear/Tanh stacks run on the CUDA path")
    layers = list(seq)
    linears = []
    for k, layer in enumerate(layers):
        want_linear = k % 2 == 0
        if want_linear and isinstance(layer, torch.nn.Linear) and layer.bias is not None:
            linears.append(layer)
        elif not want_linear and isinstance(layer, torch.nn.Tanh):
            continue
        else:
            raise NotImplementedError(f"NNProperty: unsupported layer {layer!r} at position {k}")
    if len(layers) % 2 == 0 or len(linears) < 2 or linears[-1].out_features != 1:
        raise NotImplementedError("NNProperty: expected Linear,Tanh,...,Linear(h,1)")
    width = linears[0].out_features
    for lin in linears[1:-1]:
        if lin.in_features != width or lin.out_features != width:
            raise NotImplementedError("NNProperty: hidden layers must share one width")
    if linears[-1].in_features != width:
        raise NotImplementedError("NNProperty: output layer width mismatch")
    return linears


class NNProperty(Property):
    """Property approximated by a network: ``softplus(net(x)) * scale`` (fem/properties.py:57-196)."""

    def __init__(self, net: Any, input_dim: int = 1, enforce_positive: bool = True, scale: float = 1.0):
        self.net = net
        self.input_dim = int(input_dim)
        self.enforce_positive = bool(enforce_positive)
        self.scale = float(scale)

    # -- architecture / parameters ------------------------------------------------
    @property
    def spec(self) -> ops.NetSpec:
        lin = _linear_stack(self.net)
        return ops.NetSpec(lin[0].in_features, len(lin) - 1, lin[0].out_features)

    def flat_theta(self, device) -> torch.Tensor:
        """Parameters in ``nn.Module.parameters()`` order as one fp64 vector on ``device``."""
        with torch.no_grad():
            return torch.cat([p.detach().reshape(-1).to(device=device, dtype=torch.float64)
                              for p in self.net.parameters()]).contiguous()

    def set_flat_theta(self, theta: torch.Tensor) -> None:
        off = 0
        with torch.no_grad():
            for p in self.net.parameters():
                n = p.numel()
                p.copy_(theta[off:off + n].reshape(p.shape).to(device=p.device, dtype=p.dtype))
                off += n

    # -- evaluation -----------------------------------------------------------------
    def _marshal(self, inputs) -> np.ndarray:
        """Input marshalling of fem/properties.py:113-143: dict keys are SORTED, scalars/arrays
        are padded with zeros up to ``input_dim``."""
        if inputs is None:
            return np.zeros((1, self.input_dim))
        if isinstance(inputs, dict):
            cols = [np.atleast_1d(np.asarray(inputs[k], dtype=float)) for k in sorted(inputs.keys())]
            return np.column_stack(cols)
        x = np.atleast_1d(np.asarray(inputs, dtype=float))
        if x.ndim == 1:
            x = x.reshape(1, -1) if x.size == self.input_dim else x.reshape(-1, 1)
        if x.shape[1] < self.input_dim:
            x = np.column_stack([x, np.zeros((x.shape[0], self.input_dim - x.shape[1]))])
        return x

    def evaluate(self, X, device=None) -> torch.Tensor:
        """Batched evaluation at the rows of ``X`` (``[n, input_dim]``): fp64 device tensor ``[n]``."""
        device = torch.device(device if device is not None else "cuda")
        Xd = torch.as_tensor(np.ascontiguousarray(X), dtype=torch.float64).to(device)
        spec = self.spec
        if Xd.shape[1] != spec.input_dim:
            # same failure mode as the reference, whose nn.Linear rejects the width (properties.py:116-125)
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({Xd.shape[0]}x{Xd.shape[1]} and "
                               f"{spec.input_dim}x{spec.width})")
        return ops.mlp_forward(spec, self.flat_theta(device), Xd, scale=self.scale,
                               enforce_positive=self.enforce_positive)

    def value(self, inputs: Optional[Any] = None):
        out = self.evaluate(self._marshal(inputs))
        if torch.is_grad_enabled():
            res = out.reshape(-1, 1)
            return res.squeeze() if (inputs is None or isinstance(inputs, (int, float))) else res
        arr = out.cpu().numpy()
        arr = arr.squeeze()
        if inputs is None or isinstance(inputs, (int, float)):
            return float(arr) if arr.size == 1 else arr
        return arr

    def is_trainable(self) -> bool:
        return True

    def get_torch_params(self) -> list:
        return list(self.net.parameters())

    def __repr__(self) -> str:
        n = sum(p.numel() for p in self.net.parameters())
        return f"NNProperty(dim={self.input_dim}, params={n}, scale={self.scale:.3e})"


def to_property(value: Any) -> Property:
    if isinstance(value, Property):
        return value
    if isinstance(value, (int, float)):
        return ScalarProperty(float(value))
    raise TypeError(f"Cannot convert {type(value)} to Property")
