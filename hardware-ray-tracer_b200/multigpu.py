"""Tile-parallel rendering over the GPUs of one box.

Every rank (one process per GPU, torch.distributed) holds a replica of the scene and a `brt_context`
created with (tile_rank, tile_world) = (rank, world): it traces only the 32x32 tiles with
tile_id % world == rank and packs them tile-major. One collective per frame — an all-gather of the
equal-sized packed tile buffers (NCCL over NVLink on GPUs; gloo in the CPU tests) — then the un-tile
kernel rebuilds the row-major RGBA32F frame on every rank. Smart-Culling results and the BVH are identical
on every rank by construction (same inputs, deterministic builder), so nothing else is exchanged.
"""
import torch
import torch.distributed as dist


class TiledFrame:
    """Per-rank buffers + the gather / un-tile step for frames of one size."""

    def __init__(self, ctx, width, height, rank, world, device, group=None):
        self.ctx, self.width, self.height, self.rank, self.world, self.group = ctx, width, height, rank, world, group
        n = ctx.tile_buffer_bytes(width, height, world) // 4
        self.tiles = torch.zeros(n, dtype=torch.float32, device=device)
        self.gathered = torch.zeros(n * world, dtype=torch.float32, device=device) if world > 1 else self.tiles
        self.image = torch.zeros(height * width * 4, dtype=torch.float32, device=device)

    def render(self, uniform, opts):
        """Traces this rank's tiles, gathers everybody's, returns the full (H, W, 4) frame (a view of self.image)."""
        assert opts.width == self.width and opts.height == self.height
        self.ctx.render_frame_tiles(uniform, opts, self.tiles.data_ptr())
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.tiles, group=self.group)
        self.ctx.untile(self.gathered.data_ptr(), self.width, self.height, self.world, self.image.data_ptr())
        return self.image.view(self.height, self.width, 4)
