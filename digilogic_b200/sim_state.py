"""Host mirror of the netcode `SimState` (reference crates/digilogic_netcode/src/lib.rs:169-311):
all nets of a client bit-packed LSB-first into two planes, with the invariant that the planes
always hold at least one unused bit.  Written as whole-integer arithmetic, not byte shuffling;
`tests/test_sim_state.py` pins it to the reference's 18 golden vectors."""
from __future__ import annotations


class SimState:
    def __init__(self, order: int = 0, bit_len: int = 0, bit_plane_0: bytes = b"\0", bit_plane_1: bytes = b"\0"):
        self.order = order
        self.bit_len = bit_len
        self._p0 = int.from_bytes(bit_plane_0, "little")
        self._p1 = int.from_bytes(bit_plane_1, "little")

    def reset(self, order: int = 0):
        self.order, self.bit_len, self._p0, self._p1 = order, 0, 0, 0

    def _n_bytes(self) -> int:
        return self.bit_len // 8 + 1

    @property
    def bit_plane_0(self) -> bytes:
        return self._p0.to_bytes(self._n_bytes(), "little")

    @property
    def bit_plane_1(self) -> bytes:
        return self._p1.to_bytes(self._n_bytes(), "little")

    def push_net(self, bit_width: int, bit_plane_0: bytes, bit_plane_1: bytes) -> int:
        """Appends a net; returns its bit offset (lib.rs:241-270). Bits past bit_width are masked."""
        assert 1 <= bit_width <= 255
        n = (bit_width + 7) // 8
        assert len(bit_plane_0) >= n and len(bit_plane_1) >= n
        mask = (1 << bit_width) - 1
        offset = self.bit_len
        self._p0 |= (int.from_bytes(bytes(bit_plane_0[:n]), "little") & mask) << offset
        self._p1 |= (int.from_bytes(bytes(bit_plane_1[:n]), "little") & mask) << offset
        self.bit_len += bit_width
        return offset

    def get_net(self, offset: int, bit_width: int):
        """-> (plane0 bytes, plane1 bytes) of ceil(width/8) bytes (lib.rs:272-311)."""
        assert offset + bit_width <= self.bit_len
        n = (bit_width + 7) // 8
        mask = (1 << bit_width) - 1
        return (((self._p0 >> offset) & mask).to_bytes(n, "little"), ((self._p1 >> offset) & mask).to_bytes(n, "little"))

    @classmethod
    def from_planes(cls, bit_len: int, bit_plane_0: bytes, bit_plane_1: bytes, order: int = 0) -> "SimState":
        """From the engine's bulk read-back (b2_pack_sim_state / b2_server_sim_state)."""
        s = cls(order, bit_len, bit_plane_0, bit_plane_1)
        keep = (1 << bit_len) - 1
        s._p0 &= keep
        s._p1 &= keep
        return s
