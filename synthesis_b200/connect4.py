"""Host-side mirror of the reference's `Game` trait object for Connect4 9x7
(study-connect4/src/connect4.rs:108-293): same method names and meaning, used to carry
positions across the C ABI (`ReplayBuffer.games`, `MCTS.exploit(game=...)`).  The search itself
never runs here — it runs in the CUDA engine on the two bitboards.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

WIDTH, HEIGHT = 9, 7
MAX_NUM_ACTIONS = 9
MAX_TURNS = 63
NAME = "Connect4"
NUM_PLAYERS = 2
DIMS = (1, 1, HEIGHT, WIDTH)
RED, BLACK = 0, 1  # PlayerId (connect4.rs:18-22); Red moves first

_M64 = (1 << 64) - 1
_FAB_COL = 0x7F
_FAB_ROW = sum(1 << (7 * c) for c in range(9))
_COLS05 = sum(_FAB_COL << (7 * c) for c in range(6))
_D1 = _COLS05 & sum(_FAB_ROW << r for r in (3, 4, 5, 6))
_D2 = _COLS05 & sum(_FAB_ROW << r for r in (0, 1, 2, 3))
_H = _COLS05
_V = sum(_FAB_ROW << r for r in (0, 1, 2, 3))


def won(bb: int) -> bool:  # connect4.rs:77-83
    d1 = bb & (bb >> 6) & (bb >> 12) & (bb >> 18) & _D1
    d2 = bb & (bb >> 8) & (bb >> 16) & (bb >> 24) & _D2
    h = bb & (bb >> 7) & (bb >> 14) & (bb >> 21) & _H
    v = bb & (bb >> 1) & (bb >> 2) & (bb >> 3) & _V
    return (d1 | d2 | h | v) != 0


@dataclass(eq=True)
class Connect4:
    my_bb: int = 0
    op_bb: int = 0
    height: List[int] = field(default_factory=lambda: [0] * WIDTH)
    player_: int = RED

    @staticmethod
    def new() -> "Connect4":
        return Connect4()

    @staticmethod
    def from_bitboards(my_bb: int, op_bb: int) -> "Connect4":
        occ = my_bb | op_bb
        h = [bin((occ >> (7 * c)) & _FAB_COL).count("1") for c in range(WIDTH)]
        return Connect4(int(my_bb), int(op_bb), h, sum(h) & 1)

    def clone(self) -> "Connect4":
        return Connect4(self.my_bb, self.op_bb, list(self.height), self.player_)

    def __hash__(self):  # connect4.rs:116-121: the two bitboards only
        return hash((self.my_bb, self.op_bb))

    def player(self) -> int:
        return self.player_

    def winner(self):
        return (self.player_ ^ 1) if won(self.op_bb) else None

    def is_over(self) -> bool:
        return self.winner() is not None or all(h == HEIGHT for h in self.height)

    def reward(self, player_id: int) -> float:
        w = self.winner()
        if w is None:
            return 0.0
        return 1.0 if w == player_id else -1.0

    def iter_actions(self):
        return (c for c in range(WIDTH) if self.height[c] < HEIGHT)

    def step(self, action: int) -> bool:
        col = int(action)
        if not (0 <= col < WIDTH) or self.height[col] >= HEIGHT:
            raise ValueError(f"illegal move {action}")
        self.my_bb ^= 1 << (self.height[col] + HEIGHT * col)
        self.height[col] += 1
        self.my_bb, self.op_bb = self.op_bb, self.my_bb
        self.player_ ^= 1
        return self.is_over()

    def features(self) -> np.ndarray:  # connect4.rs:237-258, shape [1][7][9]
        s = np.full((1, HEIGHT, WIDTH), -0.1, dtype=np.float32)
        for row in range(HEIGHT):
            for col in range(WIDTH):
                index = 1 << (row + HEIGHT * col)
                if self.my_bb & index:
                    s[0, row, col] = 1.0
                elif self.op_bb & index:
                    s[0, row, col] = -1.0
        for col in range(WIDTH):
            if self.height[col] < HEIGHT:
                s[0, self.height[col], col] = 0.1
        return s
