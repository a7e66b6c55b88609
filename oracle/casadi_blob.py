"""Oracle (TEST INFRASTRUCTURE): evaluate the reference's serialized CasADi
SXFunctions (bound_planner/RobotModel/*.ca) WITHOUT CasADi.

The reference ships its symbolic forward kinematics as CasADi 3.6 serialized
``SXFunction`` blobs (RobotModel.py:158,179,209,229 load them with
``ca.Function.load``): fk_pos.ca, fk_pos_col_{0..5}.ca, hom_trans.ca,
jacobian.ca.  They were generated from Pinocchio's CasADi bindings
(RobotModel.py:150-157, GEN_CA), so evaluating them reproduces the reference's
own FK numbers and pins the FK oracle / kernel against reference-generated data.

File format (CasADi serializing_stream): every byte is written as two
characters 'a'+low nibble, 'a'+high nibble.  The byte stream starts with a
magic number, holds the function header (name, input/output sparsities ...)
and ends with the SX virtual-machine program: n_instr records of
    struct { int32 op; int32 i0; union { double d; struct { int32 i1, i2; }; }; }
(constants are embedded in OP_CONST records).  Only what is needed to run the
program is decoded: the output sparsity, n_instr / worksize and the records.
Operation codes are casadi/core/calculus.hpp's ``enum Operation``.
"""
from __future__ import annotations

import math
import struct

import numpy as np

MAGIC = 123456789012345

(OP_ASSIGN, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_EXP, OP_LOG, OP_POW, OP_CONSTPOW, OP_SQRT, OP_SQ, OP_TWICE,
 OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_LT, OP_LE, OP_EQ, OP_NE, OP_NOT, OP_AND, OP_OR, OP_FLOOR,
 OP_CEIL, OP_FMOD, OP_FABS, OP_SIGN, OP_COPYSIGN, OP_IF_ELSE_ZERO, OP_ERF, OP_FMIN, OP_FMAX, OP_INV, OP_SINH,
 OP_COSH, OP_TANH, OP_ASINH, OP_ACOSH, OP_ATANH, OP_ATAN2, OP_CONST, OP_INPUT, OP_OUTPUT, OP_PARAMETER) = range(48)

_UNARY = {
    OP_ASSIGN: lambda x: x, OP_NEG: lambda x: -x, OP_EXP: math.exp, OP_LOG: math.log, OP_SQRT: math.sqrt,
    OP_SQ: lambda x: x * x, OP_TWICE: lambda x: 2.0 * x, OP_SIN: math.sin, OP_COS: math.cos, OP_TAN: math.tan,
    OP_ASIN: math.asin, OP_ACOS: math.acos, OP_ATAN: math.atan, OP_FABS: abs, OP_INV: lambda x: 1.0 / x,
    OP_SINH: math.sinh, OP_COSH: math.cosh, OP_TANH: math.tanh, OP_FLOOR: math.floor, OP_CEIL: math.ceil,
    OP_SIGN: lambda x: (x > 0) - (x < 0), OP_NOT: lambda x: float(not x),
}
_BINARY = {
    OP_ADD: lambda x, y: x + y, OP_SUB: lambda x, y: x - y, OP_MUL: lambda x, y: x * y, OP_DIV: lambda x, y: x / y,
    OP_POW: lambda x, y: x ** y, OP_CONSTPOW: lambda x, y: x ** y, OP_ATAN2: math.atan2, OP_FMIN: min, OP_FMAX: max,
    OP_LT: lambda x, y: float(x < y), OP_LE: lambda x, y: float(x <= y), OP_EQ: lambda x, y: float(x == y),
    OP_NE: lambda x, y: float(x != y), OP_AND: lambda x, y: float(bool(x) and bool(y)),
    OP_OR: lambda x, y: float(bool(x) or bool(y)), OP_COPYSIGN: math.copysign, OP_FMOD: math.fmod,
    OP_IF_ELSE_ZERO: lambda x, y: y if x else 0.0,
}


class _Dual:
    """value + tangent (forward-mode AD over the smooth operations the FK programs use)"""
    __slots__ = ("v", "t")

    def __init__(self, v, t):
        self.v, self.t = v, t


_DUAL_UNARY = {
    OP_ASSIGN: lambda x: x, OP_NEG: lambda x: _Dual(-x.v, -x.t), OP_SQ: lambda x: _Dual(x.v * x.v, 2.0 * x.v * x.t),
    OP_TWICE: lambda x: _Dual(2.0 * x.v, 2.0 * x.t), OP_SIN: lambda x: _Dual(math.sin(x.v), math.cos(x.v) * x.t),
    OP_COS: lambda x: _Dual(math.cos(x.v), -math.sin(x.v) * x.t),
    OP_SQRT: lambda x: _Dual(math.sqrt(x.v), 0.5 * x.t / math.sqrt(x.v)),
    OP_INV: lambda x: _Dual(1.0 / x.v, -x.t / (x.v * x.v)),
}
_DUAL_BINARY = {
    OP_ADD: lambda x, y: _Dual(x.v + y.v, x.t + y.t), OP_SUB: lambda x, y: _Dual(x.v - y.v, x.t - y.t),
    OP_MUL: lambda x, y: _Dual(x.v * y.v, x.t * y.v + x.v * y.t),
    OP_DIV: lambda x, y: _Dual(x.v / y.v, (x.t - x.v / y.v * y.t) / y.v),
}


def decode_bytes(text):
    text = text.strip()
    return bytes(((ord(text[2 * i]) - 97) | ((ord(text[2 * i + 1]) - 97) << 4)) for i in range(len(text) // 2))


class _Reader:
    def __init__(self, b):
        self.b, self.o = b, 0
        self.shared = []          # shared objects in definition order (only sparsities are kept)

    def i64(self):
        v = struct.unpack_from("<q", self.b, self.o)[0]
        self.o += 8
        return v

    def i32(self):
        v = struct.unpack_from("<i", self.b, self.o)[0]
        self.o += 4
        return v

    def byte(self):
        v = self.b[self.o]
        self.o += 1
        return v

    def string(self):
        n = self.i32()
        s = self.b[self.o: self.o + n].decode()
        self.o += n
        return s

    def sparsity(self):
        flag = chr(self.byte())
        if flag == "r":
            return self.shared[self.i64()]
        assert flag == "d", "unexpected shared-object flag"
        n = self.i64()
        comp = [self.i64() for _ in range(n)]
        nrow, ncol = comp[0], comp[1]
        if n == 2:                                       # dense abbreviation is not used by 3.6; keep general
            colind, row = list(range(0, nrow * ncol + 1, nrow)), [r for _ in range(ncol) for r in range(nrow)]
        else:
            colind = comp[2: 2 + ncol + 1]
            row = comp[2 + ncol + 1:]
        sp = (nrow, ncol, colind, row)
        self.shared.append(sp)
        return sp


class SXFunctionBlob:
    """A decoded ``SXFunction``: call it with the input nonzeros, get the dense output(s)."""

    def __init__(self, path):
        b = decode_bytes(open(path).read())
        r = _Reader(b)
        assert r.i64() == MAGIC, "not a CasADi serialized stream"
        self.protocol = r.i64()
        assert r.byte() == 0, "debug-decorated streams are not supported"
        assert r.byte() == 5, "not a serialized Function"
        assert chr(r.byte()) == "d" and r.byte() == 0
        assert r.string() == "SXFunction"
        r.i32()                                            # ProtoFunction version
        self.name = r.string()
        r.o += 5                                           # verbose, print_time, record_time, regularity_check, error_on_fail
        r.i32()                                            # FunctionInternal version
        for _ in range(2):                                 # is_diff_in, is_diff_out
            n = r.i64()
            r.o += n
        self.sp_in = [r.sparsity() for _ in range(r.i64())]
        self.sp_out = [r.sparsity() for _ in range(r.i64())]
        self.name_in = [r.string() for _ in range(r.i64())]
        self.name_out = [r.string() for _ in range(r.i64())]
        self._find_program(b, r.o)

    def _find_program(self, b, start):
        """SXFunction body: the symbolic inputs (shared SX nodes 'd' OP_PARAMETER name), then
        int32 version, n_instr, worksize, free_vars (empty), the operations / constants node vectors,
        default_in and finally n_instr 16-byte records.  The node vectors are skipped by scanning for
        the first offset at which n_instr consecutive records validate against worksize and the
        input / output sparsities."""
        import re

        n_in_nz = [len(sp[3]) for sp in self.sp_in]
        n_out_nz = [len(sp[3]) for sp in self.sp_out]
        ends = []
        for mt in re.finditer(rb"d/\x00{7}", b[start:]):          # 'd', OP_PARAMETER as a 64-bit int
            q = start + mt.end()
            ln = struct.unpack_from("<i", b, q)[0]
            name = b[q + 4: q + 4 + ln] if 1 <= ln <= 64 else b""
            if len(name) == ln and ln and re.fullmatch(rb"[A-Za-z_]\w*", name):
                ends.append(q + 4 + ln)
            if len(ends) == sum(n_in_nz):
                break
        assert len(ends) == sum(n_in_nz), "unexpected number of symbolic inputs"
        o = ends[-1]
        version = struct.unpack_from("<i", b, o)[0]
        n, w, f = struct.unpack_from("<qqq", b, o + 4)
        assert 1 <= version <= 3 and 1 <= w <= n and f == 0, "unexpected SXFunction header"
        for p in range(o + 28, len(b) - 16 * n + 1):
            if self._valid_program(b, p, n, w, n_in_nz, n_out_nz):
                self.n_instr, self.worksize = n, w
                self.prog = [struct.unpack_from("<ii8s", b, p + 16 * k) for k in range(n)]
                return
        raise ValueError("SX program not found")

    @staticmethod
    def _valid_program(b, p, n, w, n_in_nz, n_out_nz):
        outs = 0
        for k in range(n):
            op, i0 = struct.unpack_from("<ii", b, p + 16 * k)
            if op == OP_CONST:
                if not 0 <= i0 < w:
                    return False
                continue
            i1, i2 = struct.unpack_from("<ii", b, p + 16 * k + 8)
            if op == OP_INPUT:
                if not (0 <= i0 < w and 0 <= i1 < len(n_in_nz) and 0 <= i2 < n_in_nz[i1]):
                    return False
            elif op == OP_OUTPUT:
                if not (0 <= i0 < len(n_out_nz) and 0 <= i1 < w and 0 <= i2 < n_out_nz[i0]):
                    return False
                outs += 1
            elif op in _UNARY:
                if not (0 <= i0 < w and 0 <= i1 < w):
                    return False
            elif op in _BINARY:
                if not (0 <= i0 < w and 0 <= i1 < w and 0 <= i2 < w):
                    return False
            else:
                return False
        return outs == sum(n_out_nz)

    def jvp(self, args, dargs):
        """Forward-mode directional derivative of the program: returns (value, d value / d args . dargs) as
        dense outputs.  Used to pin djacobian_fk (RobotModel.py:233-251; the reference ships no djacobian.ca)
        to the time derivative of the reference's own jacobian.ca: dJ/dt = sum_k dJ/dq_k dq_k."""
        ins = [[_Dual(float(v), float(t)) for v, t in zip(np.asarray(a, float).reshape(-1),
                                                          np.asarray(d, float).reshape(-1))]
               for a, d in zip(args, dargs)]
        val = self._run(ins, _Dual(0.0, 0.0))
        vals = [np.vectorize(lambda e: e.v, otypes=[float])(m) for m in val]
        tans = [np.vectorize(lambda e: e.t, otypes=[float])(m) for m in val]
        return (vals[0], tans[0]) if len(vals) == 1 else (vals, tans)

    def _run(self, ins, zero):
        w = [zero] * self.worksize
        outs = [[zero] * len(sp[3]) for sp in self.sp_out]
        for op, i0, rest in self.prog:
            if op == OP_CONST:
                w[i0] = _Dual(struct.unpack("<d", rest)[0], 0.0)
                continue
            i1, i2 = struct.unpack("<ii", rest)
            if op == OP_INPUT:
                w[i0] = ins[i1][i2]
            elif op == OP_OUTPUT:
                outs[i0][i2] = w[i1]
            elif op in _DUAL_UNARY:
                w[i0] = _DUAL_UNARY[op](w[i1])
            elif op in _DUAL_BINARY:
                w[i0] = _DUAL_BINARY[op](w[i1], w[i2])
            else:
                raise NotImplementedError(f"jvp: operation code {op}")
        res = []
        for (nrow, ncol, colind, row), nz in zip(self.sp_out, outs):
            m = np.empty((nrow, ncol), dtype=object)
            m[:] = zero
            for c in range(ncol):
                for k in range(colind[c], colind[c + 1]):
                    m[row[k], c] = nz[k]
            res.append(m)
        return res

    def __call__(self, *args):
        ins = [np.asarray(a, float).reshape(-1) for a in args]
        w = [0.0] * self.worksize
        outs = [[0.0] * len(sp[3]) for sp in self.sp_out]
        for op, i0, rest in self.prog:
            if op == OP_CONST:
                w[i0] = struct.unpack("<d", rest)[0]
                continue
            i1, i2 = struct.unpack("<ii", rest)
            if op == OP_INPUT:
                w[i0] = float(ins[i1][i2])
            elif op == OP_OUTPUT:
                outs[i0][i2] = w[i1]
            elif op in _UNARY:
                w[i0] = _UNARY[op](w[i1])
            else:
                w[i0] = _BINARY[op](w[i1], w[i2])
        res = []
        for (nrow, ncol, colind, row), nz in zip(self.sp_out, outs):
            m = np.zeros((nrow, ncol))
            for c in range(ncol):
                for k in range(colind[c], colind[c + 1]):
                    m[row[k], c] = nz[k]
            res.append(m)
        return res[0] if len(res) == 1 else res
