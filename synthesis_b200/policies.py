"""Policy side of the reference API (synthesis/src/policies/traits.rs, study-connect4/src/policies.rs).

`Connect4Net` holds the weights of the reference's MLP (63-128-96-64-48-12, ReLU; names
`l_1..l_5.{weight,bias}` like the tch VarStore) and packs them into the blob the engine takes.
`RolloutPolicy` selects random-rollout leaf evaluation (synthesis/src/policies/rollout.rs).
The forward pass itself runs inside the engine's kernels; `Connect4Net.eval` / `.forward` call
the engine (there is no host implementation).
"""
import numpy as np

from . import _lib as L

LAYER_DIMS = (63, 128, 96, 64, 48, 12)


class RolloutPolicy:
    """Marker for `RolloutPolicy { rng }` (policies/rollout.rs:5-7); the per-position seed is passed
    where the reference passes the seeded StdRng."""

    def __init__(self, seed: int = 0):
        self.seed = int(seed)


class Connect4Net:
    def __init__(self, params: dict):
        self.params = {}
        for l in range(5):
            i, o = LAYER_DIMS[l], LAYER_DIMS[l + 1]
            w = np.ascontiguousarray(params[f"l_{l + 1}.weight"], dtype=np.float32)
            b = np.ascontiguousarray(params[f"l_{l + 1}.bias"], dtype=np.float32)
            if w.shape != (o, i) or b.shape != (o,):
                raise ValueError(f"l_{l + 1}: expected weight {(o, i)} and bias {(o,)}, got {w.shape} / {b.shape}")
            self.params[f"l_{l + 1}.weight"], self.params[f"l_{l + 1}.bias"] = w, b

    @staticmethod
    def new(seed: int = 0) -> "Connect4Net":
        """`NNPolicy::new(&vs)` with tch's default Linear init: W, b ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
        rng = np.random.default_rng(seed)
        p = {}
        for l in range(5):
            i, o = LAYER_DIMS[l], LAYER_DIMS[l + 1]
            bound = 1.0 / np.sqrt(i)
            p[f"l_{l + 1}.weight"] = rng.uniform(-bound, bound, size=(o, i)).astype(np.float32)
            p[f"l_{l + 1}.bias"] = rng.uniform(-bound, bound, size=(o,)).astype(np.float32)
        return Connect4Net(p)

    @staticmethod
    def from_blob(blob) -> "Connect4Net":
        blob = np.asarray(blob, dtype=np.float32).reshape(-1)
        if blob.size != L.N_WEIGHTS:
            raise ValueError(f"expected {L.N_WEIGHTS} floats, got {blob.size}")
        p, off = {}, 0
        for l in range(5):
            i, o = LAYER_DIMS[l], LAYER_DIMS[l + 1]
            p[f"l_{l + 1}.weight"] = blob[off:off + i * o].reshape(o, i).copy(); off += i * o
            p[f"l_{l + 1}.bias"] = blob[off:off + o].copy(); off += o
        return Connect4Net(p)

    @staticmethod
    def load_ot(path) -> "Connect4Net":
        """`vs.load(path)` for the VarStore of Connect4Net (alpha_zero.rs:192-194): the variables are matched by name,
        and like tch a missing or mis-shaped one is an error."""
        from .weights import read_ot
        named = read_ot(path)
        missing = [f"l_{l + 1}.{k}" for l in range(5) for k in ("weight", "bias") if f"l_{l + 1}.{k}" not in named]
        if missing:
            raise KeyError(f"{path}: cannot find {missing[0]} in the archive")
        return Connect4Net(named)

    def save_ot(self, path) -> None:
        """`vs.save(path)` (alpha_zero.rs:37, 102)."""
        from .weights import write_ot
        write_ot(path, self.params)

    def blob(self) -> np.ndarray:
        """l_1.weight, l_1.bias, ..., l_5.bias flattened (the layout syn_engine_set_weights takes)."""
        parts = []
        for l in range(5):
            parts += [self.params[f"l_{l + 1}.weight"].reshape(-1), self.params[f"l_{l + 1}.bias"]]
        out = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
        assert out.size == L.N_WEIGHTS
        return out
