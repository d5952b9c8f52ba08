"""Tile-parallel rendering over the GPUs of one box.

Every rank (one process per GPU, torch.distributed) holds a replica of the scene and a `brt_context` created with
(tile_rank, tile_world) = (rank, world): it traces only its 32x32 tiles (groups of `world` consecutive tiles, ranks rotated by the group index: csrc/render_kernels.cuh, tile_of_rank). Smart-Culling results and
the BVH are identical on every rank by construction (same inputs, deterministic builder), so only finished pixels cross GPUs.

Two exchange modes:
  "nccl"  brt_render_frame_tiles packs the rank's tiles tile-major -> all_gather_into_tensor (NCCL over NVLink; gloo in the CPU
          tests) -> brt_untile rebuilds the row-major frame on every rank.
  "p2p"   fused: the resolve kernel of brt_render_frame_peers_async stores every owned pixel straight into the RECEIVERS' gather
          images through cudaIpc-mapped peer memory (NVLink stores inside the producing kernel; RGBA32F or the 8-bit present
          format), and publishes a per-frame sequence number in their flag blocks. Receivers = rank `root` only (root_only=True,
          what the north star asks for: "gather the framebuffer") or every rank. Completion is checked ON THE DEVICE: a receiver
          enqueues a one-warp wait kernel on the stream that reads the image and a release kernel behind its read; a producer's
          resolve kernel waits on the device for the release of the frame it is about to overwrite. No collective, no host
          barrier, no stream synchronisation per frame; up to N_IMAGES frames in flight per rank (slot k <-> gather image k).
          Rule: every rank submits the same frames on the same slots in the same order; a receiver fetches or releases every frame.
          The receiver's wait kernel heads its stream for most of a frame: with CUDA's default of 8 hardware work queues other streams that share
          its queue stall behind it (C5 on 8 GPUs 1.42 -> 1.25 ms per frame with 32 queues) — the package sets CUDA_DEVICE_MAX_CONNECTIONS=32
          at import unless it is set already; import it before anything initialises CUDA.
"""
import ctypes

import torch
import torch.distributed as dist

N_IMAGES = 4  # BRT_GATHER_IMAGES (include/brt.h)


class TiledFrame:
    """Per-rank buffers + the exchange step for frames of one size."""

    def __init__(self, ctx, width, height, rank, world, device, group=None, mode="nccl", root_only=False, root=0):
        self.ctx, self.width, self.height, self.rank, self.world, self.group = ctx, width, height, rank, world, group
        self.mode = mode if world > 1 else "nccl"
        self.root_only, self.root = bool(root_only), root
        self.device = device
        if self.mode == "p2p":
            # Set-up is collective and must fail collectively: a rank that cannot map a peer's memory must not leave the others
            # waiting in the next collective, so every step is attempted on every rank and the outcome is agreed on by an all-reduce.
            err = None
            try:
                mine = ctx.gather_image_export(width, height)
            except Exception as e:  # noqa: BLE001
                mine, err = None, e
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=group)
            if err is None and all(h is not None for h in handles):
                try:
                    ctx.gather_image_open(handles)
                    ctx.gather_configure(root_only, root)
                except Exception as e:  # noqa: BLE001
                    err = e
            elif err is None:
                err = RuntimeError("a peer could not export its gather images")
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # (also the barrier: everybody's images are mapped before the first peer store)
            if int(ok.item()) == 0:
                raise RuntimeError(f"fused exchange unavailable on this box: {err or 'a peer failed to map the gather images'}")
            self._copy_stream = torch.cuda.Stream(device=device)
            self._fetch_events = {}
            self._next = 0
            self._held = None
        else:
            if device.type == "cuda":
                # brt_untile and the tile packing run on the library's stream, the all-gather on torch's current stream: make them
                # the same stream so that pack -> all-gather -> un-tile -> consumer are ordered without host synchronisation
                ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
            n = ctx.tile_buffer_bytes(width, height, world) // 4
            self.tiles = torch.zeros(n, dtype=torch.float32, device=device)
            self.gathered = torch.zeros(n * world, dtype=torch.float32, device=device) if world > 1 else self.tiles
            self.image = torch.zeros(height * width * 4, dtype=torch.float32, device=device)

    @property
    def receives(self):
        return self.mode != "p2p" or not self.root_only or self.rank == self.root

    # ---- synchronous convenience (tests, single frames) ---------------------------------------------------------------------
    def render(self, uniform, opts):
        """Traces this rank's tiles and exchanges them. nccl: returns the full (H, W, 4) frame (a view of self.image). p2p: returns
        None; on a receiver the complete frame is at frame_ptr() until the next render() (which releases it)."""
        assert opts.width == self.width and opts.height == self.height
        if self.mode == "p2p":
            if self._held is not None and self.receives:
                self.ctx.gather_release(self._held)
            slot = self._next
            self._next = (self._next + 1) % N_IMAGES
            self.ctx.render_frame_peers_async(uniform, opts, slot)
            self.ctx.frame_wait(slot)  # this rank's stores have landed
            if self.receives:
                s = torch.cuda.current_stream(self.device)
                self.ctx.gather_wait(slot, s.cuda_stream)  # ... and, on the device, everybody's
                s.synchronize()
            self._held = slot
            return None
        self.ctx.render_frame_tiles(uniform, opts, self.tiles.data_ptr())
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.tiles, group=self.group)
        self.ctx.untile(self.gathered.data_ptr(), self.width, self.height, self.world, self.image.data_ptr())
        return self.image.view(self.height, self.width, 4)

    def frame_ptr(self):
        """Device pointer of the complete row-major frame of the last render()."""
        return self.ctx.gather_image(self._held) if self.mode == "p2p" else self.image.data_ptr()

    def to_host(self, host):
        """Copies the complete frame of the last render() into `host` (pinned tensor: RGBA32F, or 4 bytes per pixel for an 8-bit format)."""
        if self.mode == "p2p":
            rt = ctypes.CDLL("libcudart.so")
            rc = rt.cudaMemcpy(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(self.frame_ptr()), ctypes.c_size_t(host.numel() * host.element_size()), 2)
            if rc != 0:
                raise RuntimeError(f"cudaMemcpy failed: {rc}")
        else:
            host.copy_(self.image, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    # ---- frames in flight (p2p) ----------------------------------------------------------------------------------------------
    def submit(self, uniform, opts, slot):
        """Enqueues the frame on slot 0 .. N_IMAGES-1 (it stores into gather image `slot` of every receiver) and returns. The slot's
        previous frame must have been fetched / released by the receivers — the resolve kernel waits for that on the device."""
        assert self.mode == "p2p" and opts.width == self.width and opts.height == self.height
        self.ctx.render_frame_peers_async(uniform, opts, slot)

    def fetch_async(self, slot, host):
        """Receiver: on the copy stream, wait (device-side) until every rank's pixels of the slot's latest frame have landed, copy the
        frame to `host` (pinned), release the image. Returns at once; wait_fetch(slot) completes it."""
        assert self.mode == "p2p" and self.receives
        with torch.cuda.stream(self._copy_stream):
            self.ctx.gather_copy_to_host(slot, host.data_ptr(), self._copy_stream.cuda_stream)
            e = torch.cuda.Event()
            e.record(self._copy_stream)
        self._fetch_events[slot] = e

    def release(self, slot):
        """Receiver that does not copy the frame out: wait for it and release it (on the copy stream)."""
        assert self.mode == "p2p" and self.receives
        self.ctx.gather_wait(slot, self._copy_stream.cuda_stream)
        self.ctx.gather_release(slot, self._copy_stream.cuda_stream)

    def wait_fetch(self, slot=None):
        if self.mode != "p2p":
            return
        for k in ([slot] if slot is not None else list(self._fetch_events)):
            e = self._fetch_events.pop(k, None)
            if e is not None:
                e.synchronize()

    def check(self):
        if self.mode == "p2p" and self.ctx.gather_timed_out():
            raise RuntimeError("fused exchange: a device-side wait timed out (a rank did not submit or release a frame)")
