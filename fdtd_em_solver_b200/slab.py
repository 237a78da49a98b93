"""y-slab decomposition and halo plan (host logic; no reference counterpart -- SURVEY.md 8(e)).

Rows (axis 0) are split into contiguous slabs, one per rank.  C-order rows are contiguous, so every
halo message is ONE contiguous row.  After each leapfrog step a rank sends

    down (to rank+1):  its last owned row of h, ex, ey        -> the receiver's halo_top row
    up   (to rank-1):  its first owned row of ex              -> the receiver's ex[row_end] slot

and the receiver recomputes hn of the halo row redundantly (sources / reflectors are global-coordinate
predicates, evaluated identically in halos).  Up/down walls exist only on the first / last slab.

The C library performs this exchange itself over NCCL (fdtd_comm_init, csrc/fdtd2d.cu: exchange());
`exchange()` below is the same plan over torch.distributed point-to-point ops, used by the gloo
(CPU) tests of the protocol and available as an alternative transport for device tensors.
"""
from dataclasses import dataclass

FIELDS = ("h", "ex", "ey")


def split_rows(rows, world, min_rows=2):
    """Even contiguous split of `rows` over `world` ranks: [(begin, end), ...]."""
    if world < 1 or rows < world * min_rows:
        raise ValueError("cannot give %d ranks at least %d rows each out of %d" % (world, min_rows, rows))
    base, extra = divmod(rows, world)
    out, r = [], 0
    for k in range(world):
        n = base + (1 if k < extra else 0)
        out.append((r, r + n))
        r += n
    return out


@dataclass(frozen=True)
class Message:
    peer: int          # rank on the other side
    field: str         # "h" | "ex" | "ey"
    row: int           # GLOBAL row index of the row that travels
    count: int         # elements (cols, or cols+1 for ey)


@dataclass(frozen=True)
class HaloPlan:
    rank: int
    world: int
    row_begin: int
    row_end: int
    cols: int

    @property
    def halo_top(self):
        return 1 if self.row_begin > 0 else 0

    @property
    def row_off(self):
        """Global row stored in local row 0."""
        return self.row_begin - self.halo_top

    @property
    def local_rows(self):
        return (self.row_end - self.row_begin) + self.halo_top + 1

    def _n(self, field):
        return self.cols + 1 if field == "ey" else self.cols

    def sends(self):
        out = []
        if self.rank + 1 < self.world:
            out += [Message(self.rank + 1, f, self.row_end - 1, self._n(f)) for f in FIELDS]
        if self.rank > 0:
            out.append(Message(self.rank - 1, "ex", self.row_begin, self.cols))
        return out

    def recvs(self):
        out = []
        if self.rank + 1 < self.world:
            out.append(Message(self.rank + 1, "ex", self.row_end, self.cols))
        if self.rank > 0:
            out += [Message(self.rank - 1, f, self.row_begin - 1, self._n(f)) for f in FIELDS]
        return out

    def bytes_per_step(self, itemsize):
        return sum(m.count for m in self.sends()) * itemsize


def plan_for(rank, world, rows, cols):
    b, e = split_rows(rows, world)[rank]
    return HaloPlan(rank, world, b, e, cols)


def exchange(plan, arrays, dist, group=None):
    """Run the plan over torch.distributed.  `arrays`: dict field -> 2-D torch tensor whose row l holds
    global row plan.row_off + l (any pitch >= count; CPU tensors for gloo, CUDA tensors for nccl)."""
    ops = []
    for m in plan.sends():
        ops.append(dist.P2POp(dist.isend, arrays[m.field][m.row - plan.row_off, :m.count].contiguous(), m.peer, group))
    landing = []
    for m in plan.recvs():
        view = arrays[m.field][m.row - plan.row_off, :m.count]
        buf = view if view.is_contiguous() else view.contiguous()
        landing.append((view, buf))
        ops.append(dist.P2POp(dist.irecv, buf, m.peer, group))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for view, buf in landing:
        if buf is not view:
            view.copy_(buf)


# ---- deep halos (fdtd_set_halo_depth): D rows of h/ex/ey stored from each neighbour, ONE exchange per K-step group ------
@dataclass(frozen=True)
class DeepPlan:
    rank: int
    world: int
    row_begin: int
    row_end: int
    rows: int
    depth: int

    @property
    def halo_top(self):
        return self.depth if self.row_begin > 0 else 0

    @property
    def halo_bot(self):
        return self.depth if self.row_end < self.rows else 0

    @property
    def row_off(self):
        return self.row_begin - self.halo_top

    @property
    def local_rows(self):
        return (self.row_end - self.row_begin) + self.halo_top + self.halo_bot + 1          # fdtd_local_rows_halo

    def blocks(self):
        """[(peer, first global row sent, first global row received)]: D rows of every field each way per interface."""
        out = []
        if self.row_end < self.rows:
            out.append((self.rank + 1, self.row_end - self.depth, self.row_end))
        if self.row_begin > 0:
            out.append((self.rank - 1, self.row_begin, self.row_begin - self.depth))
        return out


def deep_plan_for(rank, world, rows, depth):
    b, e = split_rows(rows, world, min_rows=depth)[rank]        # a slab must own at least as many rows as its halos are deep
    return DeepPlan(rank, world, b, e, rows, depth)


def exchange_deep(plan, arrays, dist, group=None):
    """The library's exchange_deep (csrc/fdtd2d.cu) over torch.distributed: `arrays` as in exchange()."""
    ops, landing = [], []
    for peer, send_row, recv_row in plan.blocks():
        for f in FIELDS:
            a = arrays[f]
            ops.append(dist.P2POp(dist.isend, a[send_row - plan.row_off:send_row - plan.row_off + plan.depth].contiguous(),
                                  peer, group))
            view = a[recv_row - plan.row_off:recv_row - plan.row_off + plan.depth]
            buf = view if view.is_contiguous() else view.contiguous()
            landing.append((view, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for view, buf in landing:
        if buf is not view:
            view.copy_(buf)
