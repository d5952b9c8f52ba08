"""Tile-parallel rendering over the GPUs of one box.

Every rank (one process per GPU, torch.distributed) holds a replica of the scene and a `brt_context`
created with (tile_rank, tile_world) = (rank, world): it traces only the 32x32 tiles with
tile_id % world == rank and packs them tile-major. One collective per frame — an all-gather of the
equal-sized packed tile buffers (NCCL over NVLink on GPUs; gloo in the CPU tests) — then the un-tile
kernel rebuilds the row-major RGBA32F frame on every rank. Smart-Culling results and the BVH are identical
on every rank by construction (same inputs, deterministic builder), so nothing else is exchanged.

Two exchange modes:
  "nccl"  brt_render_frame_tiles -> all_gather_into_tensor (NCCL over NVLink; gloo in the CPU tests) -> brt_untile
  "p2p"   brt_render_frame_peers: the resolve kernel itself stores every owned pixel into all ranks' gather images through
          cudaIpc-mapped peer memory (NVLink stores inside the producing kernel), then one barrier. No collective moves
          pixels, no un-tile pass.
"""
import ctypes

import torch
import torch.distributed as dist


class TiledFrame:
    """Per-rank buffers + the gather / un-tile step for frames of one size."""

    def __init__(self, ctx, width, height, rank, world, device, group=None, mode="nccl"):
        self.ctx, self.width, self.height, self.rank, self.world, self.group = ctx, width, height, rank, world, group
        self.mode = mode if world > 1 else "nccl"
        if self.mode == "p2p":
            mine = ctx.gather_image_export(width, height)
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=group)
            ctx.gather_image_open(handles)
            dist.barrier(group=group)
            self._barrier_token = torch.zeros(1, dtype=torch.int32, device=device)
            self._rt = ctypes.CDLL("libcudart.so")
            self._copy_stream = torch.cuda.Stream(device=device)
            self._copy_pending = False
        n = ctx.tile_buffer_bytes(width, height, world) // 4
        self.tiles = torch.zeros(n, dtype=torch.float32, device=device)
        self.gathered = torch.zeros(n * world, dtype=torch.float32, device=device) if world > 1 else self.tiles
        self.image = torch.zeros(height * width * 4, dtype=torch.float32, device=device)

    def render(self, uniform, opts):
        """Traces this rank's tiles, gathers everybody's, returns the full (H, W, 4) frame (a view of self.image)."""
        assert opts.width == self.width and opts.height == self.height
        if self.mode == "p2p":
            self.ctx.render_frame_peers(uniform, opts)               # returns when this rank's peer stores have landed
            self.wait_host()  # an asynchronous copy-out of the previous frame must be done before anybody may pass the barrier:
            #                   the frame after this one overwrites that gather image
            dist.all_reduce(self._barrier_token, group=self.group)  # barrier: everybody's have
            torch.cuda.current_stream().synchronize()
            return None  # the frame is brt_gather_image(ctx) (device memory owned by the library); see frame_ptr()
        self.ctx.render_frame_tiles(uniform, opts, self.tiles.data_ptr())
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.tiles, group=self.group)
        self.ctx.untile(self.gathered.data_ptr(), self.width, self.height, self.world, self.image.data_ptr())
        return self.image.view(self.height, self.width, 4)

    def submit(self, uniform, opts, slot):
        """p2p mode, two frames in flight: enqueues the frame on slot 0 / 1 (it stores into the gather image of that parity on every
        rank) and returns. Every rank must submit the same frames in the same order and complete() them in that order."""
        assert self.mode == "p2p" and opts.width == self.width and opts.height == self.height
        self.ctx.render_frame_peers_async(uniform, opts, slot)

    def complete(self, slot):
        """Waits for this rank's stores of the slot's frame, then the barrier that makes every rank's gather image of that frame
        complete (an asynchronous copy-out of the frame before it — the image the NEXT submit will overwrite — is finished first)."""
        self.ctx.frame_wait(slot)
        self.wait_host()
        dist.all_reduce(self._barrier_token, group=self.group)
        torch.cuda.current_stream().synchronize()

    def frame_ptr(self):
        """Device pointer of the complete row-major RGBA32F frame of the last render()."""
        return self.ctx.gather_image() if self.mode == "p2p" else self.image.data_ptr()

    def to_host_async(self, host):
        """p2p mode: starts copying the complete frame of the last render() into `host` (pinned) on a side stream and returns; the
        library alternates between two gather images, so the copy may run while the next frame is traced. wait_host() (also
        called by the next render() before its barrier) completes it."""
        assert self.mode == "p2p"
        rc = self._rt.cudaMemcpyAsync(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(self.ctx.gather_image()),
                                      ctypes.c_size_t(self.height * self.width * 16), 2, ctypes.c_void_p(self._copy_stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"cudaMemcpyAsync failed: {rc}")
        self._copy_pending = True

    def wait_host(self):
        if self.mode == "p2p" and self._copy_pending:
            self._copy_stream.synchronize()
            self._copy_pending = False

    def to_host(self, host):
        """Copies the complete frame of the last render() into `host` (a pinned float32 tensor of h*w*4 elements)."""
        if self.mode == "p2p":
            rt = ctypes.CDLL("libcudart.so")
            rc = rt.cudaMemcpy(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(self.ctx.gather_image()),
                               ctypes.c_size_t(self.height * self.width * 16), 2)  # cudaMemcpyDeviceToHost
            if rc != 0:
                raise RuntimeError(f"cudaMemcpy failed: {rc}")
        else:
            host.copy_(self.image, non_blocking=True)
            torch.cuda.current_stream().synchronize()
