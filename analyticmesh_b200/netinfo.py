"""Plain-numpy description of a network, shared by the C-ABI shim, the tests and the oracle wrapper."""
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class NetInfo:
    nodes: List[int]
    arc_table: np.ndarray          # (D, cols) int32, zero padded; row = [k, src0, tm0, ...]
    weights: List[np.ndarray]      # D+1 arrays (out, in), C-contiguous
    biases: List[np.ndarray]       # D+1 arrays (out,)
    arc_tm: List[np.ndarray]       # transforms (out, in); shape (0, 0) = identity
    dtype: np.dtype = field(default=np.dtype(np.float64))

    @property
    def state_len(self):
        return int(sum(self.nodes[1:-1]))

    @property
    def key_words(self):
        """32-bit words per key as stored by the engine: padded to a multiple of 4 (one uint4)."""
        return 4 * ((self.state_len + 127) // 128)

    @staticmethod
    def from_model(model, dtype=np.float64):
        info = model.get_info()
        dt = np.dtype(dtype)

        def arr(t):
            return np.ascontiguousarray(t.detach().cpu().double().numpy().astype(dt))

        return NetInfo(nodes=list(model.nodes),
                       arc_table=np.ascontiguousarray(info['arc_table'].cpu().numpy().astype(np.int32)),
                       weights=[arr(w) for w in info['weights']],
                       biases=[arr(b) for b in info['biases']],
                       arc_tm=[arr(t).reshape(t.shape[0], t.shape[1]) for t in info['arc_tm']],
                       dtype=dt)

    def forward(self, x):
        """float64 numpy forward pass -> (f(x), states bool (N, L)); used by tests and fixtures."""
        x0 = np.asarray(x, dtype=np.float64)
        taps = {0: x0}
        h = x0
        bits = []
        D = len(self.nodes) - 2
        for i in range(D + 1):
            y = h @ self.weights[i].astype(np.float64).T + self.biases[i].astype(np.float64)
            if i >= 1:
                row = self.arc_table[i - 1]
                for j in range(int(row[0])):
                    src, tm = int(row[1 + 2 * j]), int(row[2 + 2 * j])
                    t = self.arc_tm[tm]
                    y = y + (taps[src] if t.size == 0 else taps[src] @ t.astype(np.float64).T)
            if i < D:
                h = np.maximum(y, 0.0)
                bits.append(h > 0)
                taps[i + 1] = h
            else:
                h = y
        return h.reshape(-1), np.concatenate(bits, axis=1)
