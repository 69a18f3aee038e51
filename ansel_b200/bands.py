"""One frame over several GPUs: row bands + one gather at the end of the chain.

A band is a full-width tile of the reference's tiling engine (src/develop/tiling.c:723-1075): each rank runs
the module chain on input rows [in_y0,in_y1) with roi_in.y = in_y0, keeps rows [out_y0,out_y1) and the
finished RGBA bands are exchanged once -- all-gather (every rank ends with the frame) or gather to the
exporting rank.  No collective inside the data path for the modules built here; modules whose result needs
whole-frame statistics (wavelet thresholds, the local Laplacian pyramid) are refused in banded mode and run
as replicas instead (8e: "replicas only (first build)").

The cuts come from b200_band_plan(): on RCD's 94-row block grid when the chain starts with the demosaicer
and everything after it is pointwise -- then the banded frame is bit-identical to the untiled one -- and
tiling.c-style (overlap = sum of the modules' tiling_callback overlaps) otherwise.

torch.distributed is the plumbing (NCCL over NVLink on the GPUs; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

import ansel_b200 as ab

WHOLE_FRAME_OPS = {"bilat"}  # denoiseprofile in wavelet mode is checked per data block


class Band(C.Structure):
    """b200_band_t"""
    _fields_ = [("out_y0", C.c_int), ("out_y1", C.c_int), ("in_y0", C.c_int), ("in_y1", C.c_int)]

    def __repr__(self):
        return f"Band(out=[{self.out_y0},{self.out_y1}) in=[{self.in_y0},{self.in_y1}))"


def plan(height: int, n_bands: int, grid: int = 1, halo: int = 0, align: int = 1) -> list[Band]:
    bands = (Band * n_bands)()
    f = ab.lib().b200_band_plan
    f.argtypes = [C.c_int] * 5 + [C.POINTER(Band)]
    ab.check(f(height, n_bands, grid, halo, align, bands))
    return list(bands)


@dataclass
class Node:
    op: str            # "demosaic", "colorin", ...
    data: object       # the ctypes data block behind piece->data
    channels_in: int = 4


def needs_whole_frame(n: Node) -> bool:
    """modules whose result depends on statistics or structures of the whole frame: the local Laplacian pyramid (bilat) and the
    wavelet thresholds of the profiled denoise (both refuse tiling in the reference too)"""
    return n.op in WHOLE_FRAME_OPS or (n.op == "denoiseprofile" and n.data.mode not in (ab.DENOISE_NLMEANS, ab.DENOISE_NLMEANS_AUTO))


def chain_cuts(nodes: list[Node], width: int, height: int) -> tuple[int, int, int]:
    """(grid, halo, align) for a chain, from each module's own tiling numbers."""
    L = ab.lib()
    overlaps, aligns = [], []
    for n in nodes:
        if needs_whole_frame(n):
            raise NotImplementedError(f"{n.op}: needs whole-frame statistics -- run it as replicas or in a SegmentedChain (SURVEY.md 8e)")
        if n.op == "colorspace":
            overlaps.append(0)
            aligns.append(1)
            continue
        piece = ab.make_piece(width, height, filters=0x94949494 if n.channels_in == 1 else 0, channels=n.channels_in, data=n.data)
        t = ab.Tiling()
        getattr(L, f"b200_{n.op}_tiling")(C.byref(piece), C.byref(t))
        overlaps.append(int(t.overlap))
        aligns.append(max(1, int(t.yalign)))
    align = 1
    for a in aligns:
        align = align * a // math.gcd(align, a)
    extra = 0
    if nodes and nodes[0].op == "demosaic":
        dd = nodes[0].data
        # green equilibration FULL / BOTH averages the two greens over the buffer it is given (basic.c:296-329): every band
        # would get its own scale -- the reference's tiling has the same flaw and the module asks for no tiling then
        if int(getattr(dd, "green_eq", 0)) in (ab.GREEN_EQ_FULL, ab.GREEN_EQ_BOTH):
            raise NotImplementedError("demosaic: full-frame green equilibration needs whole-frame sums -- run it as replicas")
        # every colour-smoothing pass reads one more ring of border-treated pixels next to a cut (basic.c:192-246)
        extra = int(getattr(dd, "color_smoothing", 0))
    if nodes and nodes[0].op == "demosaic" and not any(overlaps[1:]) and not extra:
        g, h_, a_ = C.c_int(), C.c_int(), C.c_int()
        piece = ab.make_piece(width, height, filters=0x94949494, channels=1, data=nodes[0].data)
        L.b200_demosaic_band_grid(C.byref(piece), C.byref(g), C.byref(h_), C.byref(a_))
        return g.value, h_.value, a_.value
    return 1, sum(overlaps) + extra, align


def nlm_slice_height(height: int) -> int:
    """compute_slice_height(), pixel/nlmeans_core.c:267-295: the height of the chunks non-local means cuts a region of `height` rows into"""
    SLICE = 60
    if height % SLICE == 0:
        return SLICE
    best, best_incr = height % SLICE, 0
    for incr in range(1, 10):
        plus_rem = height % (SLICE + incr)
        if plus_rem == 0:
            return SLICE + incr
        if plus_rem > best:
            best_incr, best = incr, plus_rem
        minus_rem = height % (SLICE - incr)
        if minus_rem == 0:
            return SLICE - incr
        if minus_rem > best:
            best_incr, best = -incr, minus_rem
    return SLICE + best_incr


def _uses_nlm(n: Node) -> bool:
    return n.op == "nlmeans" or (n.op == "denoiseprofile" and n.data.mode in (ab.DENOISE_NLMEANS, ab.DENOISE_NLMEANS_AUTO))


def widen_for_nlm(bands: list[Band], height: int, align: int, max_chunk: int = 64) -> list[Band]:
    """Non-local means chunks the rows it is given by compute_slice_height(); chunks of up to 64 rows take the pipelined kernel
    (nlm_group.cuh), taller ones the slower group kernel.  A band's input rows may always grow into the frame (more halo than the modules
    ask for changes nothing in the rows kept), so each band takes the fewest extra rows -- in steps of the row alignment, below or above,
    wherever the frame has them -- that bring its chunk height down to 64 or less."""
    out = []
    for b in bands:
        y0, y1 = b.in_y0, b.in_y1
        if y1 > y0 and nlm_slice_height(y1 - y0) > max_chunk:
            found = None
            for extra in range(align, 4 * max_chunk + 1, align):
                for up in range(0, extra + 1, align):             # rows added on top; the rest below
                    ny0, ny1 = y0 - up, y1 + (extra - up)
                    if ny0 >= 0 and ny1 <= height and nlm_slice_height(ny1 - ny0) <= max_chunk:
                        found = (ny0, ny1)
                        break
                if found:
                    break
            if found:
                y0, y1 = found
        nb = Band()
        nb.out_y0, nb.out_y1, nb.in_y0, nb.in_y1 = b.out_y0, b.out_y1, y0, y1
        out.append(nb)
    return out


class _DeviceArray:
    """a raw device allocation seen through __cuda_array_interface__ (zero-copy torch.as_tensor)"""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _cuda_process(op, piece, src, dst, stream, node=None):
    if op == "colorspace":   # the pipe's RGB <-> Lab glue between modules (pixelpipe_hb.c dt_ioppr_transform_image_colorspace): pointwise
        cst_from, cst_to, pm = node.data
        f = ab.lib().b200_colorspace_transform_dev
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ab.ProfileMatrices), C.c_int, C.c_void_p]
        ab.check(f(src.data_ptr(), dst.data_ptr(), piece.roi_in.width, piece.roi_in.height, cst_from, cst_to, C.byref(pm), 0, stream))
        return
    ab.check(getattr(ab.lib(), f"b200_{op}_process_dev")(C.byref(piece), src.data_ptr(), dst.data_ptr(), stream))


class BandedChain:
    """Runs `nodes` on this rank's band of a width x height frame and assembles the frame.

    process: callable(op, piece, src_tensor, dst_tensor, stream) -- the CUDA library by default; the CPU tests
    of the sharding logic pass a numpy stand-in (no GPU there, and no CPU fallback in the product).
    """

    def __init__(self, nodes: list[Node], width: int, height: int, rank: int, world: int, device=None, process=None,
                 filters: int = 0x94949494, p2p: bool = False, p2p_dst=None):
        import torch
        self.torch = torch
        self.nodes, self.w, self.h, self.rank, self.world = nodes, width, height, rank, world
        self.device = device if device is not None else torch.device("cpu")
        self.process = process or _cuda_process
        self.grid, self.halo, self.align = chain_cuts(nodes, width, height)
        self.bands = plan(height, world, self.grid, self.halo, self.align)
        if self.grid == 1 and any(_uses_nlm(n) for n in nodes):
            self.bands = widen_for_nlm(self.bands, height, self.align)
        self.band = self.bands[rank]
        bh = self.band.in_y1 - self.band.in_y0
        self.pieces = []
        for n in nodes:
            p = ab.make_piece(width, max(bh, 1), filters=filters if n.channels_in == 1 else 0, channels=n.channels_in,
                              data=None if n.op == "colorspace" else n.data,
                              roi_y=self.band.in_y0, devid=self.device.index if self.device.type == "cuda" else -1)
            p.buf_in_width, p.buf_in_height = width, height  # the full frame, as tiling.c leaves piece->buf_in
            self.pieces.append(p)
        self.tmp = [torch.empty((max(bh, 1), width, 4), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.p2p = bool(p2p) and world > 1
        self.p2p_dst = p2p_dst  # None: every rank ends with the frame (all-gather); int: only that rank does (gather)
        self._own_ptr, self._peer_ptr = None, {}
        heights = [q.out_y1 - q.out_y0 for q in self.bands]
        self.equal_bands = len(set(heights)) == 1 and all(q.out_y0 == r * heights[0] for r, q in enumerate(self.bands))
        self.collective = ("one ncclAllGather of the finished RGBA bands straight into the frame (equal bands, no copy)" if self.equal_bands else
                           "one ncclAllGather of the finished RGBA bands through slots of the tallest band, then device copies into the frame")
        self.slots = None
        if not self.p2p:
            self.frame = torch.empty((height, width, 4), dtype=torch.float32, device=self.device)
            if world > 1 and not self.equal_bands:
                self.slots = torch.empty((world, max(heights), width, 4), dtype=torch.float32, device=self.device)
        else:
            self.collective = "none: the last kernel stores into the peers' frames (CUDA IPC), then one barrier"
            self._map_peer_frames()

    # ---- the gather fused into the last kernel (b200_colorout_process_scatter_dev) -----------------------------
    def _map_peer_frames(self):
        """every rank allocates its frame with b200_dev_alloc, exports it over CUDA IPC and maps the others'"""
        import torch.distributed as dist
        torch = self.torch
        if self.nodes[-1].op != "colorout":
            raise NotImplementedError("p2p gather: the chain must end with colorout (the kernel that carries the scatter)")
        L = ab.lib()
        nbytes = self.h * self.w * 16
        own = C.c_void_p()
        ab.check(L.b200_dev_alloc(C.byref(own), C.c_size_t(nbytes)))
        self._own_ptr = own.value
        self.frame = torch.as_tensor(_DeviceArray(own.value, (self.h, self.w, 4)), device=self.device)
        handle = (C.c_ubyte * 64)()
        ab.check(L.b200_ipc_export(own, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        for r, hb in enumerate(handles):
            if r == self.rank:
                continue
            p = C.c_void_p()
            ab.check(L.b200_ipc_import((C.c_ubyte * 64).from_buffer_copy(hb), C.byref(p)))
            self._peer_ptr[r] = p.value
        b = self.band
        self._last_piece = ab.make_piece(self.w, max(b.out_y1 - b.out_y0, 1), filters=0, channels=4, data=self.nodes[-1].data, roi_y=b.out_y0,
                                         devid=self.device.index)
        self._last_piece.buf_in_width, self._last_piece.buf_in_height = self.w, self.h
        row = b.out_y0 * self.w * 16
        if self.p2p_dst is None:
            ptrs = [self._own_ptr + row] + [self._peer_ptr[r] + row for r in sorted(self._peer_ptr)]
        else:  # gather to the exporting rank: one destination
            ptrs = [(self._own_ptr if self.p2p_dst == self.rank else self._peer_ptr[self.p2p_dst]) + row]
        self._dsts = (C.c_void_p * len(ptrs))(*ptrs)

    def close(self):
        """unmap the peers' frames and free this rank's (p2p mode); collective"""
        if not self.p2p or self._own_ptr is None:
            return
        import torch.distributed as dist
        L = ab.lib()
        self.torch.cuda.synchronize()
        for p in self._peer_ptr.values():
            L.b200_ipc_release(C.c_void_p(p))
        dist.barrier()                       # nobody frees a frame a peer still has mapped
        self.frame = None
        L.b200_dev_free(C.c_void_p(self._own_ptr))
        self._own_ptr, self._peer_ptr = None, {}

    def run_scatter(self, band_in, stream=0):
        """chain over the band; the last module writes the owned rows into every rank's frame itself"""
        import torch.distributed as dist
        b = self.band
        if b.out_y1 > b.out_y0:
            src = band_in
            for k, (n, p) in enumerate(zip(self.nodes[:-1], self.pieces[:-1])):
                dst = self.tmp[k & 1]
                self._run(n, p, src, dst, stream)
                src = dst
            kept = src[b.out_y0 - b.in_y0:b.out_y1 - b.in_y0]      # pointwise last module: only the owned rows
            ab.check(ab.lib().b200_colorout_process_scatter_dev(C.byref(self._last_piece), C.c_void_p(kept.data_ptr()), len(self._dsts),
                                                               self._dsts, C.c_void_p(stream)))
        # the stores of every rank must have landed before anyone reads its frame
        dist.barrier()
        return self.frame if self.p2p_dst is None or self.p2p_dst == self.rank else None

    def band_rows(self, frame_in):
        """this rank's input rows of a full-frame host/device array"""
        return frame_in[self.band.in_y0:self.band.in_y1]

    def run_band(self, band_in, stream=0):
        """chain over the band; returns a view of the rows this rank owns"""
        b = self.band
        if b.out_y1 == b.out_y0:
            return self.tmp[0][:0]
        src = band_in
        for k, (n, p) in enumerate(zip(self.nodes, self.pieces)):
            dst = self.tmp[k & 1]
            self._run(n, p, src, dst, stream)
            src = dst
        return src[b.out_y0 - b.in_y0:b.out_y1 - b.in_y0]

    def _run(self, n, p, src, dst, stream):
        if self.process is _cuda_process:
            _cuda_process(n.op, p, src, dst, stream, node=n)
        else:
            self.process(n.op, p, src, dst, stream)

    def assemble(self, mine, mode: str = "allgather", dst_rank: int = 0):
        """mode 'allgather': every rank returns the finished frame; 'gather': only dst_rank does (others None).
        ONE collective moves the pixels: ncclAllGather (dist.all_gather_into_tensor).  Bands of equal height are gathered
        straight into the frame (each rank's chain writes its band where the collective expects it: no copy at all); bands of
        unequal height (cuts on a block grid) go through slots of the tallest band's size and are copied out on the device."""
        torch = self.torch
        b = self.band
        own = self.frame[b.out_y0:b.out_y1]
        if self.world == 1:
            if b.out_y1 > b.out_y0:
                own.copy_(mine)
            return self.frame
        import torch.distributed as dist
        if mode == "allgather":
            if self.equal_bands:
                own.copy_(mine)
                dist.all_gather_into_tensor(self.frame, own)
            else:
                slot = self.slots[self.rank]
                if b.out_y1 > b.out_y0:
                    slot[:b.out_y1 - b.out_y0].copy_(mine)
                dist.all_gather_into_tensor(self.slots.view(-1, self.w, 4), slot)  # the concatenated form: rank r's slot is rows r*maxh..
                for r, q in enumerate(self.bands):
                    if q.out_y1 > q.out_y0:
                        self.frame[q.out_y0:q.out_y1].copy_(self.slots[r, :q.out_y1 - q.out_y0])
            return self.frame
        if mode == "gather":
            if b.out_y1 > b.out_y0:
                own.copy_(mine)
            ops = []
            if self.rank == dst_rank:
                ops = [dist.P2POp(dist.irecv, self.frame[q.out_y0:q.out_y1], r) for r, q in enumerate(self.bands)
                       if r != dst_rank and q.out_y1 > q.out_y0]
            elif b.out_y1 > b.out_y0:
                ops = [dist.P2POp(dist.isend, self.frame[b.out_y0:b.out_y1], dst_rank)]
            if ops:
                for wk in dist.batch_isend_irecv(ops):
                    wk.wait()
            return self.frame if self.rank == dst_rank else None
        raise ValueError(mode)

    def __call__(self, band_in, mode: str = "allgather", stream=0):
        if self.p2p:
            return self.run_scatter(band_in, stream)
        return self.assemble(self.run_band(band_in, stream), mode)


class SegmentedChain:
    """A chain with whole-frame modules in it, over several GPUs (BASELINE.json configs[3], "C4"): the nodes are cut into
    banded segments at every module that needs the whole frame (needs_whole_frame).  A banded segment runs as a BandedChain on
    this rank's rows (its own cuts: halo = its modules' overlaps); in front of a whole-frame module ONE all-gather assembles
    the frame on every rank, the module runs on the full frame on every rank (replicated: its result is the untiled one, bit for
    bit), and the next banded segment cuts its input rows out of that frame locally.  One more all-gather ends the chain.
    Collectives per frame: 1 + the number of whole-frame modules."""

    def __init__(self, nodes: list[Node], width: int, height: int, rank: int, world: int, device=None, process=None, filters: int = 0x94949494):
        import torch
        self.torch = torch
        self.w, self.h, self.rank, self.world = width, height, rank, world
        self.device = device if device is not None else torch.device("cpu")
        self.process = process or _cuda_process
        self.segments = []   # ("bands", BandedChain) | ("whole", node, piece)
        run = []
        for n in nodes:
            if needs_whole_frame(n):
                if run:
                    self.segments.append(("bands", BandedChain(run, width, height, rank, world, device=self.device, process=process, filters=filters)))
                    run = []
                p = ab.make_piece(width, height, filters=0, channels=n.channels_in, data=n.data, devid=self.device.index if self.device.type == "cuda" else -1)
                self.segments.append(("whole", n, p))
            else:
                run.append(n)
        if run:
            self.segments.append(("bands", BandedChain(run, width, height, rank, world, device=self.device, process=process, filters=filters)))
        self.full = [torch.empty((height, width, 4), dtype=torch.float32, device=self.device) for _ in range(2)] if any(s[0] == "whole" for s in self.segments) else None
        self.collectives = sum(1 for s in self.segments if s[0] == "bands")
        self.plan = [(s[0], [n.op for n in s[1].nodes], s[1].halo) if s[0] == "bands" else ("whole", [s[1].op], None) for s in self.segments]

    def band_rows(self, frame_in):
        seg = self.segments[0]
        return seg[1].band_rows(frame_in) if seg[0] == "bands" else frame_in

    def __call__(self, first_in, stream=0):
        """first_in: this rank's input rows of the first banded segment (band_rows of the frame input).  Returns the frame."""
        cur_full, cur = None, first_in
        for k, seg in enumerate(self.segments):
            if seg[0] == "bands":
                ch = seg[1]
                band_in = cur if cur_full is None else ch.band_rows(cur_full)
                cur_full = ch.assemble(ch.run_band(band_in, stream), "allgather")
                cur = None
            else:
                _kind, n, p = seg
                src = cur_full if cur_full is not None else cur
                dst = self.full[0] if src.data_ptr() != self.full[0].data_ptr() else self.full[1]
                if self.process is _cuda_process:
                    _cuda_process(n.op, p, src, dst, stream, node=n)
                else:
                    self.process(n.op, p, src, dst, stream)
                cur_full = dst
        return cur_full
