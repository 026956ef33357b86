"""Homomorphic evaluation of the PRINCE block cipher over the DHS scheme, written against the
cuHE public interface -- the caller side of BASELINE configs[3] and the ONLY fixed known-answer
the reference carries for the hot path (examples/Prince/Prince.cu:96: 64 encrypted zero bits under
k0 = 1^64, k1 = 0^64 must decrypt to 9fb51935fc3df524; per-round states at :109-144).

Test infrastructure.  The cipher itself is restated from the PRINCE specification (Borghoff et al.,
ASIACRYPT 2012): S-box table, M' built from the M0..M3 blocks, the ShiftRows nibble permutation and
the round constants are generated here, and tests/test_prince_circuit.py proves them against the
specification's test vector before any ciphertext is involved.  What follows the reference is the
*homomorphic schedule* of one S-box layer (examples/Prince/Prince.cu:191-322 and :324-455): which
products are relinearised, where modSwitch happens and in which domain each addition is done, because
that schedule is what fixes the noise budget at (d=25, p=2, w=16, min=25, cut=25, m=21845).

The evaluator is engine-agnostic: `ch` is either the cuhe_b200 module (GPU) or
tests/oracle_engine.OracleEngine (the CPU oracle behind the same interface)."""
from __future__ import annotations

from typing import List, Sequence

PRINCE_PARAMS = (25, 2, 16, 25, 25, 21845)          # examples/Prince/Prince.cu:66
KAT_HEX = "9fb51935fc3df524"                        # examples/Prince/Prince.cu:96 (PRINCE spec test vector 2)

# ---- specification constants ------------------------------------------------------------------
SBOX = [0xB, 0xF, 0x3, 0x2, 0xA, 0xC, 0x9, 0x1, 0x6, 0x7, 0x8, 0x0, 0xE, 0x5, 0xD, 0x4]
SBOX_INV = [SBOX.index(i) for i in range(16)]
RC_HEX = ["0000000000000000", "13198a2e03707344", "a4093822299f31d0", "082efa98ec4e6c89",
          "452821e638d01377", "be5466cf34e90c6c", "7ef84f78fd955cb1", "85840851f1ac43aa",
          "c882d32f25323c54", "64a51195e0e3610d", "d3b5a399ca0c2399", "c0ac29b7c97c50dd"]
SHIFT_ROWS = [0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11]   # new nibble i <- old nibble SR[i]


def hex_to_bits(h: str) -> List[int]:
    """most significant bit first, the order of the reference's bit arrays"""
    v = int(h, 16)
    return [(v >> (63 - i)) & 1 for i in range(64)]


def bits_to_hex(bits: Sequence[int]) -> str:
    v = 0
    for b in bits:
        v = (v << 1) | (int(b) & 1)
    return f"{v:016x}"


RC_BITS = [hex_to_bits(h) for h in RC_HEX]


def _m_prime_rows() -> List[List[int]]:
    """M' = diag(M^0, M^1, M^1, M^0) as, per output bit, the list of input bits that are summed."""
    def block(j):                                    # M_j = identity with diagonal entry j cleared
        return [[1 if (r == c and r != j) else 0 for c in range(4)] for r in range(4)]
    def mhat(first):
        rows = []
        for br in range(4):
            for r in range(4):
                row = []
                for bc in range(4):
                    row += block((first + br + bc) % 4)[r]
                rows.append(row)
        return rows
    out = []
    for chunk, first in enumerate((0, 1, 1, 0)):
        for row in mhat(first):
            out.append([16 * chunk + c for c, bit in enumerate(row) if bit])
    return out


M_PRIME = _m_prime_rows()


def _anf(table: Sequence[int]):
    """Algebraic normal form of a 4-bit S-box (input bits a,b,c,d = MSB..LSB) via the Moebius
    transform: for every output bit the list of monomials (tuples of input indices)."""
    outs = []
    for ob in range(4):
        f = [(table[x] >> (3 - ob)) & 1 for x in range(16)]
        for i in range(4):
            for x in range(16):
                if x & (1 << i):
                    f[x] ^= f[x ^ (1 << i)]
        monos = []
        for x in range(16):
            if f[x]:
                monos.append(tuple(j for j in range(4) if x & (8 >> j)))
        outs.append(monos)
    return outs


ANF_FWD = _anf(SBOX)
ANF_INV = _anf(SBOX_INV)


# ---- the cipher over an abstract bit type ----------------------------------------------------
class BitOps:
    """Plain bits: the circuit below evaluated in the clear (used to validate the circuit)."""

    def add(self, x, y):
        return x ^ y

    def add_const(self, x, bit):
        return x ^ bit

    def sbox_layer(self, state, inverse):
        table = SBOX_INV if inverse else SBOX
        out = []
        for i in range(16):
            v = table[int("".join(str(b) for b in state[4 * i:4 * i + 4]), 2)]
            out += [(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1]
        return out

    def finish(self, state):
        return state


class AnfBitOps(BitOps):
    """Plain bits again, but the S-box computed from its ANF with the same pair/triple products the
    homomorphic schedule forms -- proves the schedule's polynomial identities."""

    def sbox_layer(self, state, inverse):
        anf = ANF_INV if inverse else ANF_FWD
        out = []
        for i in range(16):
            x = state[4 * i:4 * i + 4]
            for monos in anf:
                v = 0
                for mono in monos:
                    t = 1
                    for j in mono:
                        t &= x[j]
                    v ^= t
                out.append(v)
        return out


def prince_eval(ops, msg, k0, k1, on_round=None):
    """PRINCE encryption of `msg` under (k0, k1) over `ops`; the order of the steps is
    examples/Prince/Prince.cu:146-189 (= the specification's)."""
    def add_key(s, k):
        return [ops.add(a, b) for a, b in zip(s, k)]

    def add_rc(s, r):
        return [ops.add_const(a, RC_BITS[r][i]) for i, a in enumerate(s)]

    def m_prime(s):
        out = []
        for srcs in M_PRIME:
            acc = s[srcs[0]]
            for j in srcs[1:]:
                acc = ops.add(acc, s[j])
            out.append(acc)
        return out

    def shift_rows(s, inverse):
        out = [None] * 64
        for i in range(16):
            src, dst = (SHIFT_ROWS[i], i) if not inverse else (i, SHIFT_ROWS[i])
            out[4 * dst:4 * dst + 4] = s[4 * src:4 * src + 4]
        return out

    s = add_key(list(msg), k0)
    s = add_key(s, k1)
    s = add_rc(s, 0)
    rnd = 0
    for _ in range(5):
        rnd += 1
        s = ops.sbox_layer(s, False)
        if on_round:
            on_round(rnd - 1, s)
        s = shift_rows(m_prime(s), False)
        s = add_rc(s, rnd)
        s = add_key(s, k1)
    s = ops.sbox_layer(s, False)
    if on_round:
        on_round(rnd, s)
    s = m_prime(s)
    s = ops.sbox_layer(s, True)
    if on_round:
        on_round(rnd + 1, s)
    for _ in range(5):
        rnd += 1
        s = add_key(s, k1)
        s = add_rc(s, rnd)
        s = m_prime(shift_rows(s, True))
        s = ops.sbox_layer(s, True)
        if on_round:
            on_round(rnd + 1, s)
    rnd += 1
    s = add_rc(s, rnd)
    s = add_key(s, k1)
    k0p = [k0[63]] + list(k0[:63])                               # k0' = (k0 >>> 1) ^ (k0 >> 63)
    k0p[63] = ops.add(k0p[63], k0[0])
    s = add_key(s, k0p)
    return ops.finish(s)


# per-round states of the KAT, after each S-box layer (examples/Prince/Prince.cu:109-144 lists the same
# twelve strings; here they are produced by the plain cipher above and compared in the test)
def kat_round_states():
    states = {}
    prince_eval(BitOps(), [0] * 64, [1] * 64, [0] * 64, on_round=lambda r, s: states.__setitem__(r, list(s)))
    return states


# ---- homomorphic evaluation ----------------------------------------------------------------------
class HomOps:
    """Ciphertext bits are ZZX values (lists of Python ints) between S-box layers, exactly as the
    reference keeps them (linear layers are host additions, Prince.cu:460-468); an S-box layer moves
    four of them to the device, evaluates the ANF there and brings four back two levels deeper."""

    def __init__(self, ch, dhs):
        self.ch, self.dhs = ch, dhs
        self.level = 0
        self.n = dhs.n
        self.counts = dict(cAnd=0, relin=0, modSwitch=0, sbox=0)

    def add(self, x, y):
        return [a + b for a, b in zip(x, y)]

    def add_const(self, x, bit):
        if not bit:
            return x
        out = list(x)
        out[0] += 1
        return out

    def finish(self, state):
        return [self.dhs.reduce(x, self.dhs.par.depth - 1) for x in state]   # Prince.cu:187-188

    def sbox_layer(self, state, inverse):
        state = [self.dhs.reduce(x, self.level) for x in state]              # Prince.cu:192-193
        out = []
        for i in range(16):
            out += self._sbox(state[4 * i:4 * i + 4], self.level, ANF_INV if inverse else ANF_FWD)
        self.level += 2
        return out

    def _sbox(self, bits, lvl, anf):
        ch = self.ch
        C = ch.CuCtxt
        x = self._inputs(bits, lvl)
        pairs = {}
        for i in range(4):
            for j in range(i + 1, 4):
                pairs[(i, j)] = C()
                ch.cAnd(pairs[(i, j)], x[i], x[j])
                self.counts["cAnd"] += 1
        # only ab and cd are multiplied again, so only they are relinearised here (Prince.cu:227-229)
        for key in ((0, 1), (2, 3)):
            pairs[key].relin()
            self.counts["relin"] += 1
        for c in list(pairs.values()) + x:
            c.modSwitch()
            self.counts["modSwitch"] += 1
        # degree <= 2 part, CRT domain, one level down
        outs = []
        for monos in anf:
            acc = None
            for mono in monos:
                if len(mono) not in (1, 2):
                    continue
                term = x[mono[0]] if len(mono) == 1 else pairs[mono]
                if acc is None:
                    acc = C()
                    ch.copy(acc, term)
                else:
                    ch.cXor(acc, acc, term)
            if () in monos:
                ch.cNot(acc, acc)
            outs.append(acc)
        # cubic terms: abd = ab*d, acd = cd*a, bcd = cd*b, abc = ab*c (Prince.cu:270-279)
        for c in x + [pairs[(0, 1)], pairs[(2, 3)]]:
            c.x2n()
        triples = {}
        for mono, (pk, single) in {(0, 1, 3): ((0, 1), 3), (0, 2, 3): ((2, 3), 0),
                                   (1, 2, 3): ((2, 3), 1), (0, 1, 2): ((0, 1), 2)}.items():
            t = C()
            ch.cAnd(t, pairs[pk], x[single])
            self.counts["cAnd"] += 1
            t.x2c()
            triples[mono] = t
        for acc, monos in zip(outs, anf):
            for mono in monos:
                if len(mono) == 3:
                    ch.cXor(acc, acc, triples[mono])
            assert not any(len(mono) == 4 for mono in monos)
        for acc in outs:
            acc.relin()
            acc.modSwitch()
            self.counts["relin"] += 1
            self.counts["modSwitch"] += 1
        for c in x + list(pairs.values()) + list(triples.values()):
            c.reset()
        self.counts["sbox"] += 1
        return self._outputs(outs)

    def _inputs(self, bits, lvl):
        """ZZX values from the host, as the reference does (Prince.cu:210-217)"""
        x = [self.ch.CuCtxt() for _ in range(4)]
        for c, v in zip(x, bits):
            c.setLevel(lvl, 0, v)
            c.x2n()
        return x

    def _outputs(self, outs):
        """back to host ZZX values (Prince.cu:315-320)"""
        res = []
        for acc in outs:
            acc.x2z()
            res.append(acc.zRep())
            acc.reset()
        return res

    def upload(self, zzx):
        return zzx

    def to_zzx(self, c):
        return c


class DeviceHomOps(HomOps):
    """the same evaluation with every ciphertext bit RESIDENT ON THE DEVICE, in the CRT
    domain, for the whole cipher.  Linear layers are cXor / cNot on the device instead of host ZZX
    additions, fresh key bits are brought to the current level with CuCtxt.dropToLevel (an extension:
    keep the first rows) instead of a host coeffReduce + re-upload, and an S-box layer consumes and
    produces device ciphertexts: no polynomial crosses PCIe between encryption and decryption (the
    reference moves 768 polynomials each way, Prince.cu:204-322)."""

    def __init__(self, ch, dhs):
        super().__init__(ch, dhs)
        self._views = {}

    def upload(self, zzx):
        c = self.ch.CuCtxt()
        c.setLevel(0, 0, zzx)
        c.x2c()
        return c

    def _at_level(self, c, lvl):
        if c.level() == lvl:
            return c
        key = (id(c), lvl)
        if key not in self._views:                      # only the long-lived key bits ever need this
            v = self.ch.CuCtxt()
            self.ch.copy(v, c)
            v.dropToLevel(lvl)
            self._views[key] = (v, c)                   # keep `c` alive so that id(c) stays unique
        return self._views[key][0]

    def add(self, x, y):
        lvl = max(x.level(), y.level())
        out = self.ch.CuCtxt()
        self.ch.cXor(out, self._at_level(x, lvl), self._at_level(y, lvl))
        return out

    def add_const(self, x, bit):
        if not bit:
            return x
        out = self.ch.CuCtxt()
        self.ch.cNot(out, x)                            # + (p - 1) = + 1 on coefficient 0 for p = 2
        return out

    def finish(self, state):
        return state

    def sbox_layer(self, state, inverse):
        out = []
        for i in range(16):
            out += self._sbox(state[4 * i:4 * i + 4], self.level, ANF_INV if inverse else ANF_FWD)
        self.level += 2
        self._views = {k: v for k, v in self._views.items() if k[1] >= self.level}
        return out

    def _inputs(self, bits, lvl):
        x = []
        for c in bits:
            t = self.ch.CuCtxt()
            self.ch.copy(t, self._at_level(c, lvl))     # the S-box transforms its inputs in place
            t.x2n()
            x.append(t)
        return x

    def _outputs(self, outs):
        return outs                                      # CRT domain, two levels down, still on the device

    def to_zzx(self, c):
        t = self.ch.CuCtxt()
        self.ch.copy(t, c)
        t.x2z()
        return t.zRep()


class BatchedHomOps(DeviceHomOps):
    """Device-resident evaluation with the 16 S-boxes of a layer evaluated TOGETHER: bit j of every nibble goes into one
    cuhe_b200.circuit.CtxtBatch of 16 ciphertexts, and the S-box schedule of the reference (which products are
    relinearised, where modSwitch happens, Prince.cu:204-322) runs once per layer on batches -- one launch set per
    operation for all 16 S-boxes instead of 16 launch sets (the reference spreads them over OpenMP threads / GPUs,
    Prince.cu:191-201).  Same operations on the same data, hence the same ciphertext words as DeviceHomOps."""

    def __init__(self, ch, dhs):
        super().__init__(ch, dhs)
        from cuhe_b200.circuit import BatchOps
        self.bops = BatchOps(0)

    def sbox_layer(self, state, inverse):
        B, lvl = self.bops, self.level
        anf = ANF_INV if inverse else ANF_FWD
        xs = [B.stack([self._at_level(state[4 * i + j], lvl) for i in range(16)]) for j in range(4)]
        outs = self._sbox16(xs, anf)
        per_bit = [B.unstack(o) for o in outs]
        self.level += 2
        self._views = {k: v for k, v in self._views.items() if k[1] >= self.level}
        self.counts["sbox"] += 16
        self.counts.update({k: B.counts[k] for k in ("cAnd", "relin", "modSwitch")})
        return [per_bit[j][i] for i in range(16) for j in range(4)]

    def _sbox16(self, xs, anf):
        B = self.bops
        xn = [B.to_ntt(x.clone()) for x in xs]                         # x2n of the four inputs
        pairs = {}
        for i in range(4):
            for j in range(i + 1, 4):
                pairs[(i, j)] = B.band(xn[i], xn[j])                   # cAnd; n2c happens inside relin / modSwitch
        for key in ((0, 1), (2, 3)):                                    # only ab and cd are multiplied again
            B.relin_(pairs[key])
        for c in list(pairs.values()) + xs:
            B.mod_switch_(c)
        outs = []
        for monos in anf:                                               # degree <= 2 part, one level down
            acc = None
            for mono in monos:
                if len(mono) not in (1, 2):
                    continue
                term = xs[mono[0]] if len(mono) == 1 else pairs[mono]
                acc = term.clone() if acc is None else B.bxor_(acc, term)
            if () in monos:
                B.bnot_(acc)
            outs.append(acc)
        xn1 = [B.to_ntt(x.clone()) for x in xs]                         # cubic terms: abd, acd, bcd, abc
        pn = {k: B.to_ntt(pairs[k].clone()) for k in ((0, 1), (2, 3))}
        triples = {}
        for mono, (pk, single) in {(0, 1, 3): ((0, 1), 3), (0, 2, 3): ((2, 3), 0),
                                   (1, 2, 3): ((2, 3), 1), (0, 1, 2): ((0, 1), 2)}.items():
            triples[mono] = B.band(pn[pk], xn1[single])
        for acc, monos in zip(outs, anf):
            for mono in monos:
                if len(mono) == 3:
                    B.bxor_(acc, triples[mono])
        for acc in outs:
            B.relin_(acc)
            B.mod_switch_(acc)
        return outs


def hom_prince(ch, dhs, msg_bits, k0_bits, k1_bits, check_rounds=(), log=None, resident=False):
    """Encrypt the 192 input bits at level 0 (Prince.cu:68-81), evaluate, decrypt at the last level
    (Prince.cu:91-94).  Returns (decrypted 64 bits, HomOps)."""
    import time as _time

    def _sync():
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        except Exception:                               # the oracle engine has no device
            pass
    ops = (BatchedHomOps(ch, dhs) if resident == "batched" else DeviceHomOps(ch, dhs)) if resident else HomOps(ch, dhs)
    t0 = _time.time()
    enc = lambda b: ops.upload(dhs.encrypt([b], 0))
    msg = [enc(b) for b in msg_bits]
    k0 = [enc(b) for b in k0_bits]
    k1 = [enc(b) for b in k1_bits]
    ops.seconds = {"encrypt_192_bits_host": _time.time() - t0, "round_checks": 0.0}
    want = kat_round_states() if check_rounds else {}
    got_rounds = {}

    def on_round(r, s):
        if log:
            log(f"round {r}: S-box layer done, level {ops.level}")
        if r in check_rounds:
            _sync()                                     # queued device work belongs to the evaluation, not to the check
            tc = _time.time()
            bits = [dhs.decrypt(ops.to_zzx(x), ops.level)[0] for x in s]
            got_rounds[r] = bits
            ops.seconds["round_checks"] += _time.time() - tc
            if log:
                log(f"round {r}: {''.join(map(str, bits))}")

    t1 = _time.time()
    out = prince_eval(ops, msg, k0, k1, on_round=on_round)
    _sync()
    ops.seconds["evaluate"] = _time.time() - t1 - ops.seconds["round_checks"]
    t2 = _time.time()
    last = dhs.par.depth - 1
    bits = [dhs.decrypt(ops.to_zzx(x), last)[0] for x in out]
    ops.seconds["decrypt_64_bits_host"] = _time.time() - t2
    ops.round_bits, ops.round_want = got_rounds, {r: want[r] for r in got_rounds}
    return bits, ops
