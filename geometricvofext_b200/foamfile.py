"""OpenFOAM on-disk formats either side of the path.

What the reference's cases and golden files are stored in, read and written without OpenFOAM:

* dictionaries (`system/fvSolution`, `controlDict`, `blockMeshDict`, the `boundary` file): the
  FoamFile header, `key value;` entries, `{}` sub-dictionaries, `()` lists (optionally sized),
  `$var` / `${var}` substitution, `//` and `/* */` comments, quoted regular-expression keys such as
  `"alpha.*"` (tutorials/test/plicVofAdvectionFoam/system/fvSolution:17-19) with OpenFOAM's lookup
  rule (exact key first, then the LAST matching pattern);
* `constant/polyMesh/{points,faces,owner,neighbour,boundary}` in ascii and binary
  (`arch "LSB;label=32|64;scalar=64"`; binary faces are a faceCompactList = offsets + labels);
* vol/surface fields: `internalField uniform v | nonuniform List<T> N (...)` in ascii and binary
  (the 13 `tutorials/test/exactSolutions/*/alpha.water.exact` files are binary volScalarFields)
  with their `boundaryField`, plus the label lists `processor*/constant/polyMesh/*ProcAddressing`.

Host-side plumbing only: nothing here computes on the path.
"""
import os
import re
from collections import OrderedDict

import numpy as np

from . import capi
from .mesh import Patch, PolyMesh

BANNER = """/*--------------------------------*- C++ -*----------------------------------*\\
| geometricvofext_b200: written in OpenFOAM file format                       |
\\*---------------------------------------------------------------------------*/
"""


class FoamFormatError(ValueError):
    pass


class FoamDict(OrderedDict):
    """A dictionary level.  Keys keep file order; quoted keys are regular expressions."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.patterns = []   # keys that were quoted in the file (regex keys), file order

    def lookup(self, key, default=None):
        """OpenFOAM's dictionary::lookup with pattern matching: exact match, else the last matching pattern."""
        if key in self and key not in self.patterns:
            return self[key]
        for p in reversed(self.patterns):
            if re.fullmatch(p, key):
                return self[p]
        if key in self:
            return self[key]
        return default


# ---------------------------------------------------------------------------------- tokenizer
_PUNCT = b"{}();[]"


class _Bin:
    """Placeholder token for a binary list payload cut out of the byte stream before tokenizing."""

    def __init__(self, array):
        self.array = array


def _strip_comments(raw):
    out = bytearray()
    i, n = 0, len(raw)
    while i < n:
        c = raw[i:i + 2]
        if c == b"//":
            j = raw.find(b"\n", i)
            i = n if j < 0 else j
        elif c == b"/*":
            j = raw.find(b"*/", i + 2)
            if j < 0:
                raise FoamFormatError("unterminated /* comment")
            i = j + 2
        elif raw[i:i + 1] == b'"':
            j = raw.find(b'"', i + 1)
            if j < 0:
                raise FoamFormatError("unterminated string")
            out += raw[i:j + 1]
            i = j + 1
        else:
            out.append(raw[i])
            i += 1
    return bytes(out)


def _tokenize(raw, bins=None):
    """bytes -> list of tokens: str words, '"quoted"' strings kept with their quotes, punctuation, _Bin objects."""
    toks = []
    i, n = 0, len(raw)
    while i < n:
        c = raw[i:i + 1]
        if c.isspace():
            i += 1
        elif c == b'"':
            j = raw.find(b'"', i + 1)
            toks.append(raw[i:j + 1].decode())
            i = j + 1
        elif c in (b"{", b"}", b"(", b")", b";", b"[", b"]"):
            toks.append(c.decode())
            i += 1
        elif raw[i:i + 2] == b"#{":       # verbatim code block of a coded function object / codeStream
            j = raw.find(b"#}", i + 2)
            if j < 0:
                raise FoamFormatError("unterminated #{ block")
            toks.append('"#{...#}"')
            i = j + 2
        elif c == b"\x00" and bins is not None and raw[i:i + 5] == b"\x00BIN\x00":
            j = raw.find(b"\x00", i + 5)
            toks.append(bins[int(raw[i + 5:j])])
            i = j + 1
        else:
            j = i
            while j < n and not raw[j:j + 1].isspace() and raw[j] not in _PUNCT and raw[j:j + 1] != b'"':
                j += 1
            toks.append(raw[i:j].decode("latin-1"))
            i = j
    return toks


def _atom(t):
    if isinstance(t, _Bin):
        return t.array
    if t and (t[0].isdigit() or t[0] in "+-."):
        try:
            return int(t)
        except ValueError:
            try:
                return float(t)
            except ValueError:
                return t
    return t


class _Parser:
    def __init__(self, toks, base_dir=None, strict=False):
        self.t, self.i, self.base_dir, self.strict = toks, 0, base_dir, strict

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def parse_dict_body(self, scopes, closing):
        d = FoamDict()
        scopes = scopes + [d]
        while True:
            tok = self.peek()
            if tok is None:
                if closing:
                    raise FoamFormatError("missing '}'")
                return d
            if tok == "}":
                if not closing:
                    raise FoamFormatError("unexpected '}'")
                self.next()
                if self.peek() == ";":    # `};` is accepted (fvSolution:33 of the reference's test case)
                    self.next()
                return d
            if tok == ";":
                self.next()
                continue
            key = self.next()
            if isinstance(key, _Bin) or key in "{}()[]":
                raise FoamFormatError("bad keyword %r" % (key,))
            if key.startswith("#"):
                self._directive(key, d, scopes)
                continue
            quoted = key.startswith('"')
            if quoted:
                key = key[1:-1]
            if key.startswith("$") and self.peek() == ";":     # `$other;` merges another dictionary
                self.next()
                src = self._resolve(key, scopes)
                if isinstance(src, dict):
                    d.update(src)
                continue
            if self.peek() == "{":
                self.next()
                val = self.parse_dict_body(scopes, True)
            else:
                vals = []
                while True:
                    tok = self.peek()
                    if tok is None:
                        raise FoamFormatError("missing ';' after %r" % key)
                    if tok == ";":
                        self.next()
                        break
                    vals.append(self.parse_value(scopes))
                val = vals[0] if len(vals) == 1 else (None if not vals else tuple(vals))
            d[key] = val
            if quoted and key not in d.patterns:
                d.patterns.append(key)

    def _directive(self, key, d, scopes):
        if key in ("#include", "#includeIfPresent", "#sinclude"):
            name = self.next()
            name = name[1:-1] if name.startswith('"') else name
            path = None
            for base in ([self.base_dir, os.path.dirname(self.base_dir)] if self.base_dir else []):   # file dir, then case dir
                if os.path.isfile(os.path.join(base, name)):
                    path = os.path.join(base, name)
                    break
            if path:
                with open(path, "rb") as f:
                    raw = f.read()
                _, off = read_header(raw)
                sub = _Parser(_tokenize(_strip_comments(raw[off:])), os.path.dirname(path)).parse_dict_body(scopes[:-1], False)
                d.update(sub)
                d.patterns += [p for p in sub.patterns if p not in d.patterns]
            elif key == "#include" and self.strict:
                raise FoamFormatError("#include: %s not found" % name)
            return
        if key in ("#includeEtc", "#inputMode", "#includeFunc", "#remove"):
            self.next()   # the argument; etc/ files are not available without an OpenFOAM installation
            return
        raise FoamFormatError("unsupported directive %s" % key)

    def _skip_block(self):
        """Skip a balanced { ... } block (#codeStream bodies)."""
        if self.next() != "{":
            raise FoamFormatError("expected '{'")
        depth = 1
        while depth:
            tok = self.next()
            if tok is None:
                raise FoamFormatError("missing '}'")
            depth += (tok == "{") - (tok == "}")

    def _eval(self, expr, scopes):
        """#eval{ ... } / #calc "..." : arithmetic on numbers and $variables."""
        import math

        def sub(m):
            v = self._resolve("$" + m.group(1), scopes)
            return repr(v) if isinstance(v, (int, float)) else str(v)
        e = re.sub(r"\$\{?([A-Za-z_][\w.]*)\}?", sub, expr)
        env = {k: getattr(math, k) for k in ("sqrt", "sin", "cos", "tan", "exp", "log", "floor", "ceil", "pow", "pi", "fabs")}
        env.update({"round": lambda x: float(math.floor(x + 0.5)), "pi": math.pi, "min": min, "max": max, "mag": abs, "sqr": lambda x: x * x})
        e = re.sub(r"\bpi\(\)", "pi", e)
        if not re.fullmatch(r"[\w\s.+\-*/(),%]*", e):
            return "#eval{%s}" % expr
        try:
            return eval(e, {"__builtins__": {}}, env)
        except Exception:
            return "#eval{%s}" % expr

    def _resolve(self, ref, scopes):
        name = ref[1:]
        if name.startswith("{") and name.endswith("}"):
            name = name[1:-1]
        name = name.lstrip(":")
        for sc in reversed(scopes):
            v = sc.lookup(name) if isinstance(sc, FoamDict) else sc.get(name)
            if v is not None:
                return v
            cur = sc
            for part in name.split("/" if "/" in name else "."):
                cur = cur.lookup(part) if isinstance(cur, FoamDict) else None
                if cur is None:
                    break
            if cur is not None:
                return cur
        if self.strict:
            raise FoamFormatError("undefined variable %s" % ref)
        return ref   # defined in an etc/ file or by a directive that is not evaluated here

    def parse_value(self, scopes):
        tok = self.next()
        if isinstance(tok, _Bin):
            return tok.array
        if tok == "(":
            return self.parse_list(scopes)
        if tok == "[":
            out = []
            while self.peek() != "]":
                out.append(_atom(self.next()))
            self.next()
            return ("dimensions", tuple(out))
        if tok == "{":
            return self.parse_dict_body(scopes, True)
        if tok.startswith('"'):
            return tok[1:-1]
        if tok.startswith("$"):
            return self._resolve(tok, scopes)
        if tok == "#eval":
            toks, depth = [], 0
            if self.next() != "{":
                raise FoamFormatError("#eval: expected '{'")
            while True:
                t = self.next()
                if t is None:
                    raise FoamFormatError("#eval: missing '}'")
                if t == "}" and depth == 0:
                    break
                depth += (t == "{") - (t == "}")
                toks.append(t)
            return self._eval(" ".join(toks), scopes)
        if tok == "#calc":
            e = self.next()
            return self._eval(e[1:-1] if e.startswith('"') else e, scopes)
        if tok == "#codeStream":
            self._skip_block()
            return "#codeStream"
        # sized list: N ( ... )  or  N { v }
        if tok.isdigit() and self.peek() == "(":
            self.next()
            lst = self.parse_list(scopes)
            return lst
        if tok.isdigit() and self.peek() == "{":
            self.next()
            v = self.parse_value(scopes)
            if self.next() != "}":
                raise FoamFormatError("bad uniform list")
            return [v] * int(tok)
        return _atom(tok)

    def parse_list(self, scopes):
        out = []
        while True:
            tok = self.peek()
            if tok is None:
                raise FoamFormatError("missing ')'")
            if tok == ")":
                self.next()
                return out
            if tok == "{":
                self.next()
                out.append(self.parse_dict_body(scopes, True))
                continue
            # `name { ... }` inside a list (blockMeshDict boundary, sampled surfaces)
            if (not isinstance(tok, _Bin)) and tok not in "()" and not tok.startswith("#") and self.i + 1 < len(self.t) and self.t[self.i + 1] == "{":
                name = self.next()
                self.next()
                out.append((name, self.parse_dict_body(scopes, True)))
                continue
            out.append(self.parse_value(scopes))


# ---------------------------------------------------------------------------------- header / binary payloads
_HDR_RE = re.compile(rb"FoamFile\s*\{(.*?)\}", re.S)
_ELEM = {"scalar": (1, "f"), "vector": (3, "f"), "sphericalTensor": (1, "f"), "symmTensor": (6, "f"), "tensor": (9, "f"),
         "label": (1, "i"), "point": (3, "f")}


def read_header(raw):
    """FoamFile header -> dict with format/class/object/label_bytes/scalar_bytes, and the offset of the body."""
    m = _HDR_RE.search(raw)
    if not m:
        return {"format": "ascii", "class": None, "label_bytes": 4, "scalar_bytes": 8}, 0
    body = _Parser(_tokenize(_strip_comments(m.group(1)))).parse_dict_body([], False)
    hdr = dict(body)
    arch = str(hdr.get("arch", "LSB;label=32;scalar=64"))
    la = re.search(r"label=(\d+)", arch)
    sc = re.search(r"scalar=(\d+)", arch)
    hdr["label_bytes"] = int(la.group(1)) // 8 if la else 4
    hdr["scalar_bytes"] = int(sc.group(1)) // 8 if sc else 8
    if "MSB" in arch:
        raise FoamFormatError("big-endian files are not supported")
    hdr.setdefault("format", "ascii")
    return hdr, m.end()


def _dtype(kind, hdr):
    return np.dtype("<f%d" % hdr["scalar_bytes"]) if kind == "f" else np.dtype("<i%d" % hdr["label_bytes"])


def _cut_binary(raw, start, hdr, top_level_elem):
    """Replace every binary list payload of a binary-format file by a placeholder; return (text, arrays).

    A payload follows `List<T> N (` inside fields, or a bare `N (` at the top level of a list file whose
    element type comes from the header class (`top_level_elem`)."""
    bins, out, i = [], bytearray(raw[:start]), start
    rx = re.compile(rb"(?:List<(\w+)>\s*)?(?<![\w.+-])(\d+)\s*\(")
    depth = 0
    while True:
        m = rx.search(raw, i)
        if not m:
            out += raw[i:]
            break
        pre = raw[i:m.start()]
        depth += pre.count(b"{") - pre.count(b"}")
        elem = m.group(1).decode() if m.group(1) else (top_level_elem if depth == 0 else None)
        n = int(m.group(2))
        # a sized ascii list such as `4(0 1 2 3)` or `value nonuniform List<scalar> 0()`
        if elem is None or elem not in _ELEM or n == 0:
            out += raw[i:m.end()]
            i = m.end()
            continue
        ncomp, kind = _ELEM[elem]
        dt = _dtype(kind, hdr)
        nbytes = n * ncomp * dt.itemsize
        payload = raw[m.end():m.end() + nbytes]
        if len(payload) != nbytes or raw[m.end() + nbytes:m.end() + nbytes + 1] != b")":
            raise FoamFormatError("binary list of %d %s: payload does not end with ')'" % (n, elem))
        arr = np.frombuffer(payload, dtype=dt).copy()
        if ncomp > 1:
            arr = arr.reshape(n, ncomp)
        out += pre + (b"List<%s> " % elem.encode() if m.group(1) else b"")
        out += b" \x00BIN\x00%d\x00 " % len(bins)
        bins.append(_Bin(arr))
        i = m.end() + nbytes + 1
    return bytes(out), bins


def parse_bytes(raw, top_level_elem=None, base_dir=None):
    """Any OpenFOAM file -> (header dict, FoamDict body | list for pure list files)."""
    hdr, off = read_header(raw)
    bins = None
    if hdr.get("format") == "binary":
        raw, bins = _cut_binary(raw, off, hdr, top_level_elem)
    body = _strip_comments(raw[off:]) if bins is None else _strip_comments_keep_bins(raw[off:])
    toks = _tokenize(body, bins)
    p = _Parser(toks, base_dir)
    # list files: [N] ( ... ) possibly repeated (faceCompactList holds two lists)
    if toks and (toks[0] == "(" or isinstance(toks[0], _Bin) or (isinstance(toks[0], str) and toks[0].isdigit() and len(toks) > 1
                                                               and (toks[1] == "(" or isinstance(toks[1], _Bin)))):
        lists = []
        while p.peek() is not None:
            lists.append(p.parse_value([]))
        return hdr, lists
    return hdr, p.parse_dict_body([], False)


def _strip_comments_keep_bins(raw):
    # binary payloads are already cut out, so comment stripping is safe on what is left
    return _strip_comments(raw)


def parse_file(path, top_level_elem=None):
    with open(path, "rb") as f:
        return parse_bytes(f.read(), top_level_elem, os.path.dirname(os.path.abspath(path)))


def read_dict(path):
    """A dictionary file (fvSolution, controlDict, blockMeshDict, ...) -> FoamDict."""
    hdr, body = parse_file(path)
    if not isinstance(body, FoamDict):
        raise FoamFormatError("%s is not a dictionary" % path)
    return body


# ---------------------------------------------------------------------------------- polyMesh
def _as_array(x, dtype, ncomp=1):
    a = np.asarray(x, dtype=dtype)
    return a.reshape(-1, ncomp) if ncomp > 1 else a.reshape(-1)


def read_points(path):
    hdr, lists = parse_file(path, "point")
    return np.ascontiguousarray(_as_array(lists[0], np.float64, 3))


def read_labels(path):
    """labelList files: owner, neighbour, cellProcAddressing, faceProcAddressing, ..."""
    hdr, lists = parse_file(path, "label")
    return _as_array(lists[0], np.int64).astype(np.int32)


def write_labels(path, labels, obj, location, fmt="binary", note=None):
    """A labelList file (cellProcAddressing ...)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_header("labelList", obj, location, fmt, note))
        f.write(_list_bytes(np.asarray(labels).astype(np.int32), fmt))


def read_faces(path):
    """faces -> (offsets[nF+1], labels): ascii faceList `N ( 4(a b c d) ... )` or binary faceCompactList."""
    hdr, lists = parse_file(path, "label")
    if hdr.get("class") == "faceCompactList" or (len(lists) == 2 and isinstance(lists[0], np.ndarray)):
        off, lab = _as_array(lists[0], np.int64), _as_array(lists[1], np.int64)
        return off.astype(np.int32), lab.astype(np.int32)
    faces = lists[0]
    sizes = np.fromiter((len(f) for f in faces), dtype=np.int64, count=len(faces))
    off = np.zeros(len(faces) + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    lab = np.fromiter((v for f in faces for v in f), dtype=np.int64, count=int(off[-1]))
    return off.astype(np.int32), lab.astype(np.int32)


def read_boundary(path):
    """polyBoundaryMesh: list of (name, FoamDict(type, nFaces, startFace, ...))."""
    hdr, lists = parse_file(path)
    out = []
    for item in lists[0]:
        if not (isinstance(item, tuple) and len(item) == 2):
            raise FoamFormatError("boundary: expected `name { ... }` entries")
        out.append(item)
    return out


def read_polymesh(case_dir, region=None):
    """constant/polyMesh (or processorN/constant/polyMesh when case_dir is a processor directory) -> PolyMesh."""
    d = os.path.join(case_dir, "constant", region or "", "polyMesh")
    if not os.path.isdir(d):
        d = case_dir
    points = read_points(os.path.join(d, "points"))
    off, lab = read_faces(os.path.join(d, "faces"))
    owner = read_labels(os.path.join(d, "owner"))
    neighbour = read_labels(os.path.join(d, "neighbour"))
    patches = []
    for name, pd in read_boundary(os.path.join(d, "boundary")):
        kind, nbr = capi.PATCH_GENERIC, -1
        ptype = str(pd.get("type", "patch"))
        if ptype == "empty":
            kind = capi.PATCH_EMPTY
        elif ptype in ("processor", "processorCyclic"):
            kind, nbr = capi.PATCH_PROCESSOR, int(pd["neighbProcNo"])
        patches.append(Patch(name, int(pd["startFace"]), int(pd["nFaces"]), kind=kind, nbr_rank=nbr))
    n_cells = int(max(owner.max(), neighbour.max() if neighbour.size else -1)) + 1
    m = PolyMesh(points=points, face_offsets=off, face_points=lab, owner=owner, neighbour=neighbour, patches=patches,
                 n_cells=n_cells, meta={"kind": "polyMesh", "dir": d,
                                        "patch_types": {n: str(pd.get("type", "patch")) for n, pd in read_boundary(os.path.join(d, "boundary"))}})
    return m


def _header(cls, obj, location, fmt, note=None):
    s = BANNER + "FoamFile\n{\n    version     2.0;\n    format      %s;\n" % fmt
    if fmt == "binary":
        s += '    arch        "LSB;label=32;scalar=64";\n'
    if note:
        s += '    note        "%s";\n' % note
    s += "    class       %s;\n" % cls
    if location:
        s += '    location    "%s";\n' % location
    s += "    object      %s;\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n" % obj
    return s.encode()


def _list_bytes(a, fmt, ncomp=1, prec=17):
    a = np.ascontiguousarray(a)
    n = a.shape[0]
    if fmt == "binary":
        dt = "<f8" if a.dtype.kind == "f" else "<i4"
        return b"%d\n(" % n + a.astype(dt).tobytes() + b")\n"
    if ncomp == 1:
        if a.dtype.kind == "f":
            body = "\n".join(repr(float(v)) if prec >= 17 else ("%.*g" % (prec, v)) for v in a)
        else:
            body = "\n".join(str(int(v)) for v in a)
    else:
        body = "\n".join("(" + " ".join(repr(float(v)) for v in row) + ")" for row in a)
    return ("%d\n(\n%s\n)\n" % (n, body)).encode()


def write_polymesh(mesh, case_dir, fmt="binary", patch_types=None):
    """Write constant/polyMesh/{points,faces,owner,neighbour,boundary}."""
    d = os.path.join(case_dir, "constant", "polyMesh")
    os.makedirs(d, exist_ok=True)
    note = "nPoints:%d  nCells:%d  nFaces:%d  nInternalFaces:%d" % (mesh.n_points, mesh.n_cells, mesh.n_faces, mesh.n_internal_faces)
    with open(os.path.join(d, "points"), "wb") as f:
        f.write(_header("vectorField", "points", "constant/polyMesh", fmt))
        f.write(_list_bytes(mesh.points.reshape(-1, 3), fmt, 3))
    with open(os.path.join(d, "faces"), "wb") as f:
        if fmt == "binary":
            f.write(_header("faceCompactList", "faces", "constant/polyMesh", fmt))
            f.write(_list_bytes(mesh.face_offsets.astype(np.int32), fmt))
            f.write(_list_bytes(mesh.face_points.astype(np.int32), fmt))
        else:
            f.write(_header("faceList", "faces", "constant/polyMesh", fmt))
            off, lab = mesh.face_offsets, mesh.face_points
            rows = ["%d(%s)" % (off[i + 1] - off[i], " ".join(str(int(v)) for v in lab[off[i]:off[i + 1]])) for i in range(mesh.n_faces)]
            f.write(("%d\n(\n%s\n)\n" % (mesh.n_faces, "\n".join(rows))).encode())
    for name, arr in (("owner", mesh.owner), ("neighbour", mesh.neighbour)):
        with open(os.path.join(d, name), "wb") as f:
            f.write(_header("labelList", name, "constant/polyMesh", fmt, note))
            f.write(_list_bytes(arr.astype(np.int32), fmt))
    ptypes = dict(mesh.meta.get("patch_types", {}))
    ptypes.update(patch_types or {})
    with open(os.path.join(d, "boundary"), "wb") as f:
        f.write(_header("polyBoundaryMesh", "boundary", "constant/polyMesh", "ascii"))
        s = "%d\n(\n" % len(mesh.patches)
        for p in mesh.patches:
            t = ptypes.get(p.name, "empty" if p.kind == capi.PATCH_EMPTY else "processor" if p.kind == capi.PATCH_PROCESSOR else "patch")
            s += "    %s\n    {\n        type            %s;\n" % (p.name, t)
            if t == "wall":
                s += "        inGroups        1(wall);\n"
            s += "        nFaces          %d;\n        startFace       %d;\n" % (p.size, p.start)
            if p.kind == capi.PATCH_PROCESSOR:
                s += "        matchTolerance  0.0001;\n        myProcNo        %d;\n        neighbProcNo    %d;\n" % (
                    int(mesh.meta.get("rank", 0)), p.nbr_rank)
            s += "    }\n"
        f.write((s + ")\n").encode())
    return d


# ---------------------------------------------------------------------------------- fields
_CLASS_ELEM = {"volScalarField": "scalar", "surfaceScalarField": "scalar", "volVectorField": "vector",
               "surfaceVectorField": "vector", "volScalarField::Internal": "scalar", "pointScalarField": "scalar",
               "pointVectorField": "vector"}


class FoamField:
    """internalField (scalar/tuple for `uniform`, ndarray for `nonuniform`) + boundaryField FoamDict."""

    def __init__(self, cls, name, dimensions, internal, boundary, header=None):
        self.cls, self.name, self.dimensions, self.internal, self.boundary, self.header = cls, name, dimensions, internal, boundary, header

    def internal_array(self, n):
        """Expanded internal field for n cells/faces."""
        ncomp = _ELEM[_CLASS_ELEM.get(self.cls, "scalar")][0]
        if isinstance(self.internal, np.ndarray):
            a = self.internal
            if a.shape[0] != n:
                raise FoamFormatError("%s: internalField has %d entries, mesh has %d" % (self.name, a.shape[0], n))
            return np.ascontiguousarray(a, dtype=np.float64)
        v = np.asarray(self.internal, dtype=np.float64).reshape(-1)
        return np.ascontiguousarray(np.broadcast_to(v if ncomp > 1 else v[0], (n, ncomp) if ncomp > 1 else (n,)), dtype=np.float64)

    def patch_entry(self, patch_name):
        return self.boundary.lookup(patch_name)


def _field_value(v, ncomp):
    """('uniform', x) | ('nonuniform', 'List<T>', data) token tuples -> python scalar/tuple or ndarray."""
    if isinstance(v, tuple) and len(v) >= 2 and v[0] == "uniform":
        x = v[1]
        return tuple(float(c) for c in x) if isinstance(x, list) else float(x)
    if isinstance(v, tuple) and len(v) >= 2 and v[0] == "nonuniform":
        data = v[-1]
        if isinstance(data, np.ndarray):
            return data
        if isinstance(data, (int, float)):     # `nonuniform List<scalar> 0()` tokenizes as ... 0, []
            data = []
        a = np.asarray(data, dtype=np.float64)
        return a.reshape(-1, ncomp) if ncomp > 1 else a.reshape(-1)
    raise FoamFormatError("unsupported field value %r" % (v,))


def read_field(path):
    """A vol*/surface* field file -> FoamField."""
    hdr, body = parse_file(path)
    cls = str(hdr.get("class"))
    ncomp = _ELEM[_CLASS_ELEM.get(cls, "scalar")][0]
    if "internalField" not in body:
        raise FoamFormatError("%s: no internalField" % path)
    internal = _field_value(body["internalField"], ncomp)
    bf = body.get("boundaryField", FoamDict())
    for pd in bf.values():
        if isinstance(pd, dict):
            for k in ("value", "inletValue", "refValue", "gradient"):
                if k in pd and isinstance(pd[k], tuple) and pd[k] and pd[k][0] in ("uniform", "nonuniform"):
                    pd[k] = _field_value(pd[k], ncomp)
    dims = body.get("dimensions")
    return FoamField(cls, str(hdr.get("object", os.path.basename(path))), dims[1] if isinstance(dims, tuple) else None, internal, bf, hdr)


def write_field(path, cls, name, internal, boundary, dimensions=(0, 0, 0, 0, 0, 0, 0), fmt="binary", location=None):
    """Write a vol/surface field.  boundary: {patchName: {"type": ..., "value": scalar|ndarray, ...}} in patch order."""
    elem = _CLASS_ELEM[cls]
    ncomp = _ELEM[elem][0]

    def val(v):
        if isinstance(v, np.ndarray) and (v.ndim == 2 or (ncomp == 1 and v.ndim == 1)):
            a = np.ascontiguousarray(v, dtype=np.float64)
            return b"nonuniform List<%s> " % elem.encode() + _list_bytes(a, fmt, ncomp).rstrip(b"\n")
        if ncomp == 1:
            return ("uniform %s" % repr(float(v))).encode()
        return ("uniform (%s)" % " ".join(repr(float(c)) for c in np.asarray(v).reshape(-1))).encode()

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_header(cls, name, location, fmt))
        f.write(("dimensions      [%s];\n\n" % " ".join(str(d) for d in dimensions)).encode())
        f.write(b"internalField   " + val(internal) + b";\n\nboundaryField\n{\n")
        for pname, pd in boundary.items():
            f.write(("    %s\n    {\n" % pname).encode())
            for k, v in pd.items():
                if k == "type" or isinstance(v, str):
                    f.write(("        %-15s %s;\n" % (k, v)).encode())
                else:
                    f.write(("        %-15s " % k).encode() + val(v) + b";\n")
            f.write(b"    }\n")
        f.write(b"}\n\n// ************************************************************************* //\n")


def apply_alpha_boundary(mesh, field):
    """Map the boundaryField of an alpha field onto the patches' svof boundary conditions
    (zeroGradient / fixedValue / inletOutlet are what the path implements; empty and processor follow the patch)."""
    for p in mesh.patches:
        if p.kind in (capi.PATCH_EMPTY, capi.PATCH_PROCESSOR):
            continue
        pd = field.patch_entry(p.name)
        if pd is None:
            raise FoamFormatError("field %s has no boundaryField entry for patch %s" % (field.name, p.name))
        t = str(pd.get("type"))
        if t in ("zeroGradient", "empty", "symmetry", "symmetryPlane", "constantAlphaContactAngle"):
            # contact-angle patches evaluate as zeroGradient for the transported value (gradient correction is the caller's)
            p.alpha_bc, p.alpha_value = capi.BC_ZERO_GRADIENT, 0.0
        elif t == "fixedValue":
            v = pd.get("value")
            if isinstance(v, np.ndarray):
                if v.size and np.ptp(v) != 0.0:
                    raise FoamFormatError("patch %s: non-uniform fixedValue is not supported" % p.name)
                v = float(v.flat[0]) if v.size else 0.0
            p.alpha_bc, p.alpha_value = capi.BC_FIXED_VALUE, float(v)
        elif t == "inletOutlet":
            v = pd.get("inletValue", 0.0)
            if isinstance(v, np.ndarray):
                v = float(v.flat[0]) if v.size else 0.0
            p.alpha_bc, p.alpha_value = capi.BC_INLET_OUTLET, float(v)
        else:
            raise FoamFormatError("patch %s: alpha boundary type '%s' is not supported (zeroGradient, fixedValue, inletOutlet)" % (p.name, t))
    return mesh


def write_vtk_polydata(path, points, face_offsets, cell_data=None, title="PLIC interface"):
    """Legacy-VTK polydata (ascii): the polygons of SolveVofEqu.interface() with per-polygon data
    (the reference's sampler writes the same surface as .vtp, controlDict `surfaceFormat vtp`)."""
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    off = np.asarray(face_offsets, dtype=np.int64)
    nF = off.shape[0] - 1
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 3.0\n%s\nASCII\nDATASET POLYDATA\nPOINTS %d double\n" % (title, points.shape[0]))
        for p in points:
            f.write("%.17g %.17g %.17g\n" % (p[0], p[1], p[2]))
        f.write("POLYGONS %d %d\n" % (nF, nF + int(off[-1]) if nF else 0))
        for i in range(nF):
            f.write("%d %s\n" % (off[i + 1] - off[i], " ".join(str(v) for v in range(off[i], off[i + 1]))))
        if cell_data and nF:
            f.write("CELL_DATA %d\n" % nF)
            for name, arr in cell_data.items():
                arr = np.asarray(arr)
                kind = "int" if arr.dtype.kind in "iu" else "double"
                f.write("SCALARS %s %s 1\nLOOKUP_TABLE default\n" % (name, kind))
                f.write("\n".join(("%d" % v) if kind == "int" else ("%.17g" % v) for v in arr) + "\n")
    return path
