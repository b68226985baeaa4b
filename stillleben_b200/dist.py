"""Multi-GPU plumbing: one process per GPU, scenes sharded by index, shared assets broadcast once.

Scenes of a batch are independent (the reference treats GPUs as independent processes:
python/src/py_context.cpp:12-52), so the only communication is ONE broadcast of the packed
read-only asset arena (consolidated vertex / index buffers, material tables, texture images, and the
light maps WITH their precomputed IBL maps: environment cube, irradiance, prefilter, BRDF LUT) from
rank 0 at load time — NCCL over NVLink on the GPU box, gloo in the CPU tests.  There is no
collective in the steady state.
"""
import json
import os

import numpy as np

from . import abi
from .desc import ImageData, LightMapData, MaterialData, MeshData


def shard_range(n_items, rank, world_size):
    """Contiguous block of scene indices owned by `rank` (blocks differ by at most one item)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes_weighted(n_items, weights):
    """Shard sizes proportional to `weights` (largest-remainder rounding, every rank with a positive weight gets at least one
    item when there are enough). For the host-buffer path on a box whose GPUs do not see the same device->host bandwidth
    (PCIe switches shared by different numbers of GPUs): a shard's wall time is its bytes over its link, so equal TIME per
    rank means sizes proportional to the links."""
    w = [max(float(x), 0.0) for x in weights]
    total = sum(w)
    if total <= 0.0:
        w, total = [1.0] * len(w), float(len(w))
    exact = [n_items * x / total for x in w]
    sizes = [int(e) for e in exact]
    order = sorted(range(len(w)), key=lambda i: (exact[i] - sizes[i], w[i]), reverse=True)
    for i in order[:n_items - sum(sizes)]:
        sizes[i] += 1
    return sizes


def shard_range_weighted(n_items, rank, weights):
    """Contiguous block of scene indices owned by `rank` when the blocks are sized by shard_sizes_weighted()."""
    sizes = shard_sizes_weighted(n_items, weights)
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# ---- asset arena ------------------------------------------------------------------------------
def pack_meshes(meshes, light_maps=()):
    """Serialise a list of MeshData (and LightMapData, with their precomputed maps when `maps` is set) into
    (header bytes, flat uint8 arena)."""
    chunks, metas, off = [], [], 0

    def put(a):
        nonlocal off
        a = np.ascontiguousarray(a)
        b = a.view(np.uint8).reshape(-1)
        pad = (-len(b)) % 256
        chunks.append(b)
        if pad:
            chunks.append(np.zeros(pad, np.uint8))
        at = off
        off += len(b) + pad
        return at, len(b)

    for m in meshes:
        meta = {"name": m.name, "submeshes": [tuple(int(x) for x in s) for s in m.submeshes],
                "bbox_min": np.asarray(m.bbox_min, np.float32).tolist(), "bbox_max": np.asarray(m.bbox_max, np.float32).tolist(),
                "vertices": put(m.vertices), "n_vertices": len(m.vertices), "indices": put(m.indices),
                "materials": [vars(x).copy() for x in m.materials], "images": []}
        for im in m.images:
            meta["images"].append({"shape": list(im.pixels.shape), "data": put(im.pixels), "wrap_s": im.wrap_s,
                                   "wrap_t": im.wrap_t, "min_filter": im.min_filter, "mag_filter": im.mag_filter,
                                   "kind": im.kind})
        metas.append(meta)
    lms = []
    for lm in light_maps:
        eq = np.ascontiguousarray(lm.equirect, np.float32)
        meta = {"equirect": put(eq), "equirect_shape": list(eq.shape), "light_directions": [list(map(float, d)) for d in lm.light_directions],
                "light_colors": [list(map(float, c)) for c in lm.light_colors], "maps": None}
        if lm.maps is not None:
            meta["maps"] = [{"data": put(np.ascontiguousarray(a, np.float32)), "shape": list(np.shape(a))} for a in lm.maps]
        lms.append(meta)
    arena = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return json.dumps({"meshes": metas, "light_maps": lms}).encode(), arena


def unpack_assets(header, arena):
    """-> (list of MeshData, list of LightMapData)."""
    doc = json.loads(bytes(header).decode())
    lms = []
    for meta in doc["light_maps"]:
        o, n = meta["equirect"]
        eq = arena[o:o + n].view(np.float32).reshape(meta["equirect_shape"]).copy()
        maps = None
        if meta["maps"] is not None:
            maps = tuple(arena[m["data"][0]:m["data"][0] + m["data"][1]].view(np.float32).reshape(m["shape"]).copy() for m in meta["maps"])
        lms.append(LightMapData(eq, meta["light_directions"], meta["light_colors"], maps))
    return unpack_meshes(header, arena), lms


def unpack_meshes(header, arena):
    metas = json.loads(bytes(header).decode())["meshes"]
    out = []
    for meta in metas:
        vo, vn = meta["vertices"]
        verts = arena[vo:vo + vn].view(abi.VERTEX_DTYPE).copy()
        io_, in_ = meta["indices"]
        idx = arena[io_:io_ + in_].view(np.uint32).copy()
        images = []
        for im in meta["images"]:
            o, n = im["data"]
            images.append(ImageData(arena[o:o + n].reshape(im["shape"]).copy(), im["wrap_s"], im["wrap_t"], im["min_filter"],
                                    im["mag_filter"], im["kind"]))
        mats = [MaterialData(**{k: tuple(v) if isinstance(v, list) else v for k, v in m.items()}) for m in meta["materials"]]
        out.append(MeshData(verts, idx, [tuple(s) for s in meta["submeshes"]], mats, images,
                            np.array(meta["bbox_min"], np.float32), np.array(meta["bbox_max"], np.float32), meta["name"]))
    return out


def broadcast_meshes(meshes, src=0, device=None):
    """Broadcast the mesh pool from rank `src` to every rank (one collective for the arena)."""
    return broadcast_assets(meshes, (), src, device)[0]


def precompute_light_map(ctx, lm):
    """Run the IBL precompute of `lm` on this rank's GPU (if it has not run yet) and attach the maps to it, so that the
    asset broadcast carries them and the other ranks skip the precompute (north_star: "one NCCL broadcast of shared
    mesh/IBL buffers at load time")."""
    if lm.maps is None:
        lm.maps = tuple(ctx.read_lightmap(lm))
    return lm


def broadcast_assets(meshes, light_maps=(), src=0, device=None):
    """Broadcast the mesh pool and the light maps from rank `src` to every rank: ONE collective for the whole arena.

    `meshes` / `light_maps` are only read on `src`; other ranks may pass None.  Works with any initialised
    torch.distributed backend; with NCCL the arena travels GPU-to-GPU over NVLink. Returns (meshes, light_maps)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return meshes, list(light_maps or ())
    rank = dist.get_rank()
    if rank == src:
        header, arena = pack_meshes(meshes, light_maps or ())
        sizes = torch.tensor([len(header), len(arena)], dtype=torch.int64)
    else:
        header, arena = b"", None
        sizes = torch.zeros(2, dtype=torch.int64)
    dev = device if device is not None else torch.device("cpu")
    sizes = sizes.to(dev)
    dist.broadcast(sizes, src)
    hn, an = int(sizes[0]), int(sizes[1])
    payload = torch.empty(hn + an, dtype=torch.uint8, device=dev)
    if rank == src:
        payload[:hn] = torch.frombuffer(bytearray(header), dtype=torch.uint8).to(dev)
        payload[hn:] = torch.from_numpy(arena).to(dev)
    dist.broadcast(payload, src)
    if rank == src:
        return meshes, list(light_maps or ())
    host = payload.cpu().numpy()
    return unpack_assets(host[:hn].tobytes(), host[hn:])


# ---- host placement ---------------------------------------------------------------------------
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated.

    With one process per GPU the descriptor marshalling and — more importantly — the page-locked result buffers of
    `slb_render_batch_host` should live next to the GPU: first-touch places them on the node of the allocating
    CPU, and a device->host copy into the other socket's memory crosses the inter-socket link that all remote
    GPUs share. Returns the NUMA node used, or None when the topology cannot be read (then nothing changes)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        node_path = f"/sys/bus/pci/devices/{bus}/numa_node"
        node = int(open(node_path).read().strip())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
