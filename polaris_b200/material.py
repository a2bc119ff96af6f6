"""Layered-material expressions -> flat 64-byte MaterialNode records.

Mirrors the part of the reference that turns `mat_expr` strings into the node list the
tracer uploads: the grammar of asset/material/material_expr.y:75-176, the type codes of
asset/material/bxdf.go:6-17 and op.go:7-17 (they must equal the constants the device code
switches on, CL/bxdf/bxdf.cl:13-18 and CL/samplers/material_sampler.cl:4-8), the defaults
of asset/material/defaults.go:6-13 and the post-order flattening of
asset/compiler/compiler.go:332-503.

Textures are supplied in memory (name -> (format, width, height, bytes)); the reference
loads them with OpenImageIO (asset/texure/texture.go) which is outside this path.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

# bxdf.go:6-17 (1 << iota, first entry is "invalid")
BXDF_EMISSIVE = 2
BXDF_DIFFUSE = 4
BXDF_CONDUCTOR = 8
BXDF_ROUGH_CONDUCTOR = 16
BXDF_DIELECTRIC = 32
BXDF_ROUGH_DIELECTRIC = 64
# op.go:7-17
OP_MIX = 10001
OP_MIX_MAP = 10002
OP_BUMP_MAP = 10003
OP_NORMAL_MAP = 10004
OP_DISPERSE = 10005

BXDF_BY_NAME = {
    "emissive": BXDF_EMISSIVE,
    "diffuse": BXDF_DIFFUSE,
    "conductor": BXDF_CONDUCTOR,
    "roughConductor": BXDF_ROUGH_CONDUCTOR,
    "dielectric": BXDF_DIELECTRIC,
    "roughDielectric": BXDF_ROUGH_DIELECTRIC,
}

# texture formats, asset/texure/texture_fmt.go:5-10
TEX_LUMINANCE8, TEX_LUMINANCE32F, TEX_RGBA8, TEX_RGBA32F = 0, 1, 2, 3

# A few entries of asset/material/ior.go (the table has ~245 names); lookups are
# case-insensitive (ior.go:263-277).
KNOWN_IORS = {
    "AIR": 1.0002926,
    "GLASS": 1.51714,
    "GOLD": 0.47,
    "SILVER": 0.18,
    "COPPER": 1.10,
    "DIAMOND": 2.417,
    "WATER": 1.33157,
}
DEFAULT_INT_IOR = KNOWN_IORS["GLASS"]  # defaults.go:12
DEFAULT_EXT_IOR = KNOWN_IORS["AIR"]    # defaults.go:13

MATERIAL_NODE_DTYPE = np.dtype(
    [
        ("union1", np.int32, 4),    # type, left, right|transmittanceTex, bump|mixWeights|refl|spec|radiance tex
        ("union2", np.float32, 4),  # reflectance|specularity|radiance|intDispersionIORs|mixWeight
        ("union3", np.float32, 4),  # transmittance|extDispersionIORs
        ("union4", np.float32, 3),  # intIOR, extIOR, roughness|scale
        ("union5", np.int32, 1),    # roughnessTex
    ]
)
assert MATERIAL_NODE_DTYPE.itemsize == 64  # optimized_scene.go:82-110


def align4(value: int) -> int:
    """compiler.go:556-563: next multiple of 4."""
    return value + ((-value) % 4)


class MaterialError(ValueError):
    pass


# ----------------------------------------------------------------------------- parser
_TOKEN = re.compile(r'\s*(?:(?P<num>[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)|(?P<id>[A-Za-z_]\w*)|"(?P<str>[^"]*)"|(?P<p>[(){}:,]))')


def _tokenize(s: str):
    pos, out = 0, []
    while pos < len(s):
        if s[pos:].strip() == "":
            break
        m = _TOKEN.match(s, pos)
        if not m:
            raise MaterialError(f"material expression: unexpected input at {s[pos:pos + 12]!r}")
        pos = m.end()
        if m.group("num") is not None:
            out.append(("num", np.float32(m.group("num"))))
        elif m.group("id") is not None:
            out.append(("id", m.group("id")))
        elif m.group("str") is not None:
            out.append(("str", m.group("str")))
        else:
            out.append((m.group("p"), m.group("p")))
    return out


@dataclass
class Expr:
    kind: str  # bxdf name | mix | mixMap | bumpMap | normalMap | disperse | ref
    params: dict = field(default_factory=dict)
    children: list = field(default_factory=list)


class _Parser:
    def __init__(self, text):
        self.t = _tokenize(text)
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", None)

    def take(self, kind):
        k, v = self.peek()
        if k != kind:
            raise MaterialError(f"material expression: expected {kind}, got {k} {v!r}")
        self.i += 1
        return v

    def float3(self):
        self.take("{")
        a = self.take("num"); self.take(",")
        b = self.take("num"); self.take(",")
        c = self.take("num"); self.take("}")
        return np.array([a, b, c], dtype=np.float32)

    def expr(self, top=False):
        k, v = self.peek()
        if k == "str":
            if top:
                raise MaterialError("a material reference cannot be the whole expression")
            self.i += 1
            return Expr("ref", {"name": v})
        name = self.take("id")
        self.take("(")
        if name in BXDF_BY_NAME:
            e = Expr(name)
            while self.peek()[0] != ")":
                pname = self.take("id"); self.take(":")
                if pname in ("reflectance", "specularity", "transmittance", "radiance"):
                    val = ("tex", self.take("str")) if self.peek()[0] == "str" else ("vec3", self.float3())
                elif pname in ("intIOR", "extIOR"):
                    val = ("name", self.take("str")) if self.peek()[0] == "str" else ("float", self.take("num"))
                elif pname == "scale":
                    val = ("float", self.take("num"))
                elif pname == "roughness":
                    val = ("tex", self.take("str")) if self.peek()[0] == "str" else ("float", self.take("num"))
                else:
                    raise MaterialError(f"unknown bxdf parameter {pname!r}")
                # BxdfParamNode.Validate (asset/material/node.go:138-163)
                if val[0] == "vec3":
                    comps = [float(c) for c in val[1]]
                    if pname == "reflectance" and any(c >= 1.0 for c in comps):
                        raise MaterialError(f"energy conservation violation for Parameter {pname!r}; ensure that all vector components are < 1.0")
                    if pname in ("specularity", "transmittance") and any(c > 1.0 for c in comps):
                        raise MaterialError(f"energy conservation violation for Parameter {pname!r}; ensure that all vector components are <= 1.0")
                if pname == "roughness" and val[0] == "float" and float(val[1]) > 1.0:
                    raise MaterialError(f"values for Parameter {pname!r} must be in the [0, 1] range")
                e.params[pname] = val
                if self.peek()[0] == ",":
                    self.i += 1
            self.take(")")
            return e
        if name == "mix":
            a = self.expr(); self.take(",")
            b = self.expr(); self.take(",")
            w = self.take("num"); self.take(")")
            if not (0.0 <= float(w) <= 1.0):  # node.go:220-237
                raise MaterialError("mix weight must be in [0, 1]")
            return Expr("mix", {"weight": w}, [a, b])
        if name == "mixMap":
            a = self.expr(); self.take(",")
            b = self.expr(); self.take(",")
            tex = self.take("str"); self.take(")")
            return Expr("mixMap", {"tex": tex}, [a, b])
        if name in ("bumpMap", "normalMap"):
            a = self.expr(); self.take(",")
            tex = self.take("str"); self.take(")")
            return Expr(name, {"tex": tex}, [a])
        if name == "disperse":
            a = self.expr(); self.take(",")
            k1 = self.take("id"); self.take(":"); v1 = self.float3(); self.take(",")
            k2 = self.take("id"); self.take(":"); v2 = self.float3(); self.take(")")
            if (k1, k2) != ("intIOR", "extIOR"):
                raise MaterialError("disperse expects intIOR: {..}, extIOR: {..}")
            return Expr("disperse", {"intIOR": v1, "extIOR": v2}, [a])
        raise MaterialError(f"unknown material function {name!r}")


def parse_expression(text: str) -> Expr:
    p = _Parser(text)
    e = p.expr(top=True)
    if p.peek()[0] != "eof":
        raise MaterialError("trailing input after material expression")
    return e


# ----------------------------------------------------------------------------- compiler
class MaterialCompiler:
    """compiler.go:271-553: material trees, texture baking, emissive lookup."""

    def __init__(self, materials: dict, textures: dict | None = None):
        # materials: ordered name -> expression string (used ones first, like the wavefront
        # reader leaves them, wavefront.go:192-244)
        self.materials = dict(materials)
        self.textures = textures or {}
        self.nodes = []  # list of MATERIAL_NODE_DTYPE scalars
        self.tex_meta = []  # (format, w, h, offset)
        self.tex_data = bytearray()
        self.tex_index = {}
        self._ref_stack = []

    def _bake_texture(self, name) -> int:
        if name in self.tex_index:
            return self.tex_index[name]
        if name not in self.textures:
            return -1  # "skipping missing texture" (compiler.go:510-513)
        fmt, w, h, data = self.textures[name]
        data = bytes(data)
        off = len(self.tex_data)
        self.tex_data += data
        self.tex_data += b"\0" * (align4(len(data)) - len(data))  # compiler.go:528-537
        self.tex_meta.append((fmt, w, h, off))
        self.tex_index[name] = len(self.tex_meta) - 1
        return self.tex_index[name]

    def _ior(self, val):
        kind, v = val
        if kind == "float":
            return np.float32(v)
        key = v.upper()
        if key not in KNOWN_IORS:
            raise MaterialError(f"unknown material name {v!r}; try specifying the IOR manually")
        return np.float32(KNOWN_IORS[key])

    def _gen(self, e: Expr) -> int:
        n = np.zeros((), dtype=MATERIAL_NODE_DTYPE)
        n["union1"] = (0, -1, -1, -1)
        n["union5"] = (-1,)
        n["union4"] = (DEFAULT_INT_IOR, DEFAULT_EXT_IOR, 0.0)  # on every node (compiler.go:336-341)
        if e.kind == "ref":
            name = e.params["name"]
            if name in self._ref_stack:
                raise MaterialError(f"detected circular dependency loop via {name!r}")
            if name not in self.materials:
                raise MaterialError(f"reference to undefined material {name!r}")
            return self.generate(name)
        if e.kind in BXDF_BY_NAME:
            t = BXDF_BY_NAME[e.kind]
            n["union1"][0] = t
            if t == BXDF_DIFFUSE:
                n["union2"] = (0.2, 0.2, 0.2, 0.0)
            elif t in (BXDF_CONDUCTOR, BXDF_ROUGH_CONDUCTOR):
                n["union2"] = (1, 1, 1, 0)
            elif t in (BXDF_DIELECTRIC, BXDF_ROUGH_DIELECTRIC):
                n["union2"] = (1, 1, 1, 0)
                n["union3"] = (1, 1, 1, 0)
            elif t == BXDF_EMISSIVE:
                n["union2"] = (1, 1, 1, 0)
                n["union4"][2] = 1.0
            if t in (BXDF_ROUGH_CONDUCTOR, BXDF_ROUGH_DIELECTRIC):
                n["union4"][2] = 0.1
            for pname, val in e.params.items():  # setMaterialNodeParameter (compiler.go:462-503)
                if pname in ("reflectance", "specularity", "radiance"):
                    if val[0] == "vec3":
                        n["union2"] = (*val[1], 0.0)
                    else:
                        n["union1"][3] = self._bake_texture(val[1])
                elif pname == "transmittance":
                    if val[0] == "vec3":
                        n["union3"] = (*val[1], 0.0)
                    else:
                        n["union1"][2] = self._bake_texture(val[1])
                elif pname == "intIOR":
                    n["union4"][0] = self._ior(val)
                elif pname == "extIOR":
                    n["union4"][1] = self._ior(val)
                elif pname == "scale":
                    n["union4"][2] = val[1]
                elif pname == "roughness":
                    if val[0] == "float":
                        n["union4"][2] = val[1]
                    else:
                        n["union5"][0] = self._bake_texture(val[1])
        elif e.kind == "mix":
            n["union1"][0] = OP_MIX
            n["union1"][1] = self._gen(e.children[0])
            n["union1"][2] = self._gen(e.children[1])
            n["union2"][0] = e.params["weight"]
        elif e.kind == "mixMap":
            n["union1"][0] = OP_MIX_MAP
            n["union1"][1] = self._gen(e.children[0])
            n["union1"][2] = self._gen(e.children[1])
            n["union1"][3] = self._bake_texture(e.params["tex"])
        elif e.kind in ("bumpMap", "normalMap"):
            n["union1"][0] = OP_BUMP_MAP if e.kind == "bumpMap" else OP_NORMAL_MAP
            n["union1"][1] = self._gen(e.children[0])
            n["union1"][3] = self._bake_texture(e.params["tex"])
        elif e.kind == "disperse":
            n["union1"][0] = OP_DISPERSE
            n["union1"][1] = self._gen(e.children[0])
            n["union2"] = (*e.params["intIOR"], 0.0)
            n["union3"] = (*e.params["extIOR"], 0.0)
        else:
            raise MaterialError(f"unsupported node {e.kind!r}")
        self.nodes.append(n)  # post-order: children first (compiler.go:458-459)
        return len(self.nodes) - 1

    def generate(self, name: str) -> int:
        self._ref_stack.append(name)
        try:
            return self._gen(parse_expression(self.materials[name]))
        finally:
            self._ref_stack.pop()

    def find_emissive(self, idx: int) -> int:
        """findMaterialNodeByBxdf(.., BxdfEmissive), compiler.go:244-268."""
        t = int(self.nodes[idx]["union1"][0])
        if 1 < t < 128:  # IsBxdfType
            return idx if t == BXDF_EMISSIVE else -1
        out = self.find_emissive(int(self.nodes[idx]["union1"][1]))
        if out != -1:
            return out
        if t == OP_MIX:
            out = self.find_emissive(int(self.nodes[idx]["union1"][2]))
        return out

    def node_array(self):
        if not self.nodes:
            return np.zeros(0, dtype=MATERIAL_NODE_DTYPE)
        return np.array(self.nodes, dtype=MATERIAL_NODE_DTYPE)
