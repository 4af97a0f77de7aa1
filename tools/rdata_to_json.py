#!/usr/bin/env python3
"""Convert the reference's shipped .RData inputs into JSON fixtures (tests/golden/).

Run once in the build container (where /root/reference exists); the output is
committed so nothing under tests/ or bench.py needs /root/reference at run time.

    python tools/rdata_to_json.py /root/reference/data tests/golden

Format notes (R serialisation, XDR, versions 2 and 3): gzip -> "RDX2\\n"/"RDX3\\n" ->
"X\\n" -> int version, int writer, int min_reader [v3: int nenc, bytes enc] -> one
item (a pairlist of the saved symbols).  Only the SEXP types the two data files
use are implemented.
"""
import gzip
import bz2
import lzma
import json
import struct
import sys
import os


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64(self):
        v = struct.unpack_from(">d", self.b, self.o)[0]
        self.o += 8
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if ty == 254:  # NILVALUE
            return None
        if ty == 253:  # GLOBALENV
            return {"_env": "global"}
        if ty == 242:  # EMPTYENV
            return {"_env": "empty"}
        if ty == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if ty == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if ty == 2:  # LISTSXP (pairlist)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                ty2 = flags & 0xFF
                if ty2 == 254:
                    break
                if ty2 != 2:
                    raise ValueError("unexpected pairlist tail type %d" % ty2)
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return {"_pairlist": out}
        if ty == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", "replace")
        if ty in (10, 13):  # LGLSXP / INTSXP
            n = self.i32()
            vals = [self.i32() for _ in range(n)]
            vals = [None if v == -2147483648 else v for v in vals]
            return self._with_attr({"_t": "int", "v": vals}, has_attr)
        if ty == 14:  # REALSXP
            n = self.i32()
            vals = list(struct.unpack_from(">%dd" % n, self.b, self.o))
            raw = self.b[self.o:self.o + 8 * n]
            self.o += 8 * n
            # R's NA_real_ is a NaN with low word 1954; keep both as None
            out = []
            for k, v in enumerate(vals):
                out.append(None if v != v else v)
            return self._with_attr({"_t": "real", "v": out}, has_attr)
        if ty == 16:  # STRSXP
            n = self.i32()
            vals = [self.item() for _ in range(n)]
            return self._with_attr({"_t": "str", "v": vals}, has_attr)
        if ty == 19:  # VECSXP
            n = self.i32()
            vals = [self.item() for _ in range(n)]
            return self._with_attr({"_t": "list", "v": vals}, has_attr)
        if ty == 238:  # ALTREP
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = info["_pairlist"][0][1] if isinstance(info, dict) else None
            if cls == "compact_intseq":
                n, start, step = [int(x) for x in state["v"]]
                return {"_t": "int", "v": [start + k * step for k in range(n)]}
            if cls == "compact_realseq":
                n, start, step = state["v"]
                return {"_t": "real", "v": [start + k * step for k in range(int(n))]}
            if cls in ("wrap_real", "wrap_integer", "wrap_string", "wrap_logical"):
                inner = state["v"][0] if state.get("_t") == "list" else state["_pairlist"][0][1]
                if attr is not None and isinstance(inner, dict):
                    inner = dict(inner)
                    inner["attr"] = _pl_to_dict(attr)
                return inner
            if cls == "deferred_string":
                src = state["_pairlist"][0][1]
                return {"_t": "str", "v": [repr(v) for v in src["v"]]}
            raise ValueError("unsupported ALTREP class %r" % cls)
        raise ValueError("unsupported SEXP type %d at offset %d" % (ty, self.o))

    def _with_attr(self, obj, has_attr):
        if has_attr:
            obj["attr"] = _pl_to_dict(self.item())
        return obj


def _pl_to_dict(pl):
    if pl is None:
        return {}
    return {tag: car for tag, car in pl["_pairlist"]}


def load_rdata(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    assert raw[:3] == b"RDX", raw[:8]
    assert raw[5:7] == b"X\n", raw[:8]
    r = _Reader(raw)
    r.o = 7
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()
    return _pl_to_dict(top)


def names_of(obj):
    a = obj.get("attr", {})
    nm = a.get("names")
    return nm["v"] if nm else None


def experiment_to_json(e):
    """e: VECSXP with names y, x, drug_names, experiment_ID (some optional)."""
    nm = names_of(e)
    if nm is None:  # mathews_DLBCL: unnamed list(y, x); drug names live in x's dimnames
        nm = ["y", "x"][:len(e["v"])]
    d = dict(zip(nm, e["v"]))
    y = d["y"]
    x = d["x"]
    ydim = y.get("attr", {}).get("dim")
    xdim = x.get("attr", {}).get("dim")
    nrow = xdim["v"][0] if xdim else len(x["v"]) // 2
    ycol = ydim["v"][1] if ydim else 1
    out = {
        "nrow": nrow,
        "y_ncol": ycol,
        # column-major, exactly as stored by R
        "y": y["v"],
        "x": x["v"],
    }
    dn = x.get("attr", {}).get("dimnames")
    if dn is not None and dn["v"][1] is not None:
        out["x_colnames"] = dn["v"][1]["v"]
    for k in ("drug_names", "experiment_ID", "units"):
        if k in d and d[k] is not None:
            out[k] = d[k]["v"]
    return out


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    m = load_rdata(os.path.join(src, "mathews_DLBCL.RData"))
    obj = m["mathews_DLBCL"]
    out = {}
    for name, e in zip(names_of(obj), obj["v"]):
        out[name] = experiment_to_json(e)
    with open(os.path.join(dst, "mathews_DLBCL.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("mathews_DLBCL:", {k: (v["nrow"], v["y_ncol"]) for k, v in out.items()})

    o = load_rdata(os.path.join(src, "ONeil_A375.RData"))
    obj = o["ONeil_A375"]
    d = dict(zip(names_of(obj), obj["v"]))
    failed = d["failed"]
    out = []
    for e in failed["v"]:
        out.append(experiment_to_json(e))
    with open(os.path.join(dst, "ONeil_A375_failed.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("ONeil_A375 failed:", [(v["nrow"], v["y_ncol"], v.get("drug_names")) for v in out])
    # the fitted screenSummary (a data.frame) -> columns; VB-derived, used only as a loose anchor
    ss = d["screenSummary"]
    cols = {}
    for name, col in zip(names_of(ss), ss["v"]):
        if col.get("_t") == "int" and "levels" in col.get("attr", {}):
            lv = col["attr"]["levels"]["v"]
            cols[name] = [None if v is None else lv[v - 1] for v in col["v"]]
        else:
            cols[name] = col["v"]
    with open(os.path.join(dst, "ONeil_A375_screenSummary.json"), "w") as f:
        json.dump(cols, f)
    print("screenSummary columns:", list(cols.keys()), "rows", len(next(iter(cols.values()))))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
