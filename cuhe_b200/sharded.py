"""Residue-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

The reference has no collective (one OpenMP thread per GPU, whole
ciphertexts per device).  Here the CRT-residue axis is sharded: rank r of G owns
primes r, r+G, r+2G, ... (so every level stays balanced while modSwitch drops
the last prime).  Collectives appear only where the algorithm exchanges data:

  * exchange of cRep before ICRT (cuhe/CuHE.cu:366-382 needs every residue of a coefficient):
    ICRT is split by coefficient range, so an all-to-all delivers to rank j just slice j of every
    residue (`exchange_for_icrt`; `all_gather_residues` is the simple all-gather form), then the
    RAW slices are all-gathered;
  * broadcast of the dropped residue row in modSwitch (cuhe/Base.cu:1112-1138).

The index arithmetic is pure torch (CPU or CUDA tensors) so it is tested on CPU
with the gloo backend (tests/test_dist_cpu.py); the compute goes through the C ABI.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def local_primes(L: int, rank: int, world: int) -> List[int]:
    """Prime indices owned by `rank` at a level with L primes."""
    return list(range(rank, L, world))


def rows_of(L: int, rank: int, world: int) -> int:
    return (L - rank + world - 1) // world if rank < L else 0


def owner_of(prime: int, world: int) -> Tuple[int, int]:
    """(rank, local row) holding prime index `prime`."""
    return prime % world, prime // world


def all_gather_residues(local: torch.Tensor, L: int, world: int, group=None) -> torch.Tensor:
    """local: [batch][rows][H] (rows padded to ceil(L/world) on every rank)
    -> [batch][L][H] in prime order on every rank."""
    B, rows_pad, H = local.shape
    if world == 1:
        return local[:, :L]
    gathered = torch.empty((world * B, rows_pad, H), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)    # rank-major concatenation
    gathered = gathered.view(world, B, rows_pad, H)
    # gathered[r, b, i, :] is prime r + world*i
    full = gathered.permute(1, 2, 0, 3).reshape(B, rows_pad * world, H)
    return full[:, :L].contiguous()


def exchange_for_icrt(local: torch.Tensor, L: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-to-all form of the exchange before ICRT.  ICRT on rank j only needs the coefficient slice
    [j*H/G, (j+1)*H/G) of every residue, so each rank sends slice j of its own rows to rank j
    (bytes received per rank = 1/G of what the all-gather moves).
    local: [batch][rows_pad][H]  ->  [batch][L][H/G] in prime order (this rank's coefficient slice)."""
    B, rows_pad, H = local.shape
    if world == 1:
        return local[:, :L].contiguous()
    assert H % world == 0
    Hs = H // world
    # send[j] = slice j of my rows: [world][B][rows_pad][Hs]
    send = local.view(B, rows_pad, world, Hs).permute(2, 0, 1, 3).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(world * B, rows_pad, Hs), send.view(world * B, rows_pad, Hs), group=group)
    # recv[r, b, i, :] = prime r + world*i, my slice
    return recv.permute(1, 2, 0, 3).reshape(B, rows_pad * world, Hs)[:, :L].contiguous()


def all_gather_raw_slices(mine: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """mine: [batch][H/G][W] = this rank's coefficient slice of the RAW result -> [batch][H][W]."""
    if world == 1:
        return mine
    B, Hs, W = mine.shape
    gathered = torch.empty((world * B, Hs, W), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine.contiguous(), group=group)
    return gathered.view(world, B, Hs, W).permute(1, 0, 2, 3).reshape(B, world * Hs, W).contiguous()


def all_gather_operands(mine: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """Data-parallel input: rank r uploaded products [r*b, (r+1)*b) -> every rank gets all world*b RAW
    operands ([b][H][W] -> [world*b][H][W], product order) over NVLink, so each polynomial crosses
    PCIe once in the whole job instead of once per rank."""
    if world == 1:
        return mine
    out = torch.empty((world * mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine.contiguous(), group=group)
    return out


def raw_slices_to_owners(mine: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """mine: [world*b][H/G][W] = this rank's coefficient slice of every product -> [b][H][W], the
    complete RAW result of the products this rank owns (products r*b.. on rank r).  An all-to-all:
    1/world of the traffic of all_gather_raw_slices when every rank only reads its own products back."""
    if world == 1:
        return mine
    B, Hs, W = mine.shape
    assert B % world == 0
    b = B // world
    send = mine.contiguous()                                       # [owner][b][Hs][W]
    recv = torch.empty_like(send)                                  # [source rank = slice][b][Hs][W]
    dist.all_to_all_single(recv.view(world * b, Hs, W), send.view(world * b, Hs, W), group=group)
    return recv.view(world, b, Hs, W).permute(1, 0, 2, 3).reshape(b, world * Hs, W).contiguous()


def coefficient_slice(H: int, rank: int, world: int) -> Tuple[int, int]:
    step = (H + world - 1) // world
    return min(rank * step, H), min((rank + 1) * step, H)


def all_gather_raw(raw: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """raw: [batch][H][W] with only this rank's coefficient slice valid ->
    complete RAW polynomials on every rank."""
    if world == 1:
        return raw
    B, H, W = raw.shape
    step = (H + world - 1) // world
    assert step * world == H, "crtLen is a power of two; world must divide it"
    b, e = coefficient_slice(H, rank, world)
    mine = raw[:, b:e].contiguous()
    gathered = torch.empty((world * B, step, W), dtype=raw.dtype, device=raw.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return gathered.view(world, B, step, W).permute(1, 0, 2, 3).reshape(B, H, W).contiguous()


def broadcast_last_row(local: torch.Tensor, L: int, rank: int, world: int, group=None) -> torch.Tensor:
    """local: [rows][H] residues of this rank at a level with L primes.  Returns
    the row of prime L-1 on every rank (input of cuhe_mod_switch)."""
    H = local.shape[-1]
    owner, row = owner_of(L - 1, world)
    buf = local[row].clone() if rank == owner else torch.empty(H, dtype=local.dtype, device=local.device)
    if world > 1:
        dist.broadcast(buf, src=owner, group=group)
    return buf
