"""Minimal reader / writer of Uintah data archives (UDA) in python, for the tests: an independent second implementation of the format
uintah_b200/host/DataArchive.h reads and writes (index.xml, tNNNNN/timestep.xml, tNNNNN/lL/pNNNNN.{xml,data}; gzip blocks =
8-byte little-endian size + zlib stream)."""
import os
import re
import struct
import zlib

import numpy as np

_TYPES = {"int": np.int32, "float": np.float32, "double": np.float64}


def _iv(s):
    return tuple(int(round(float(x))) for x in re.sub(r"[\[\],]", " ", s).split())


def _vv(s):
    return tuple(float(x) for x in re.sub(r"[\[\],]", " ", s).split())


def read_grid(uda, tdir):
    """[(level dict)] with anchor, dx, extra, patches [(id, lo, hi, interior_lo, interior_hi, proc)]; <Grid> from grid.xml when present"""
    gx = os.path.join(uda, tdir, "grid.xml")
    x = open(gx if os.path.exists(gx) else os.path.join(uda, tdir, "timestep.xml")).read()
    levels = []
    for lev in re.findall(r"<Level>(.*?)</Level>", x, re.S):
        g = lambda tag, src=lev: re.search(rf"<{tag}>([^<]*)</{tag}>", src).group(1)
        patches = []
        for p in re.findall(r"<Patch>(.*?)</Patch>", lev, re.S):
            lo, hi = _iv(g("lowIndex", p)), _iv(g("highIndex", p))
            m = re.search(r"<interiorLowIndex>([^<]*)<", p)
            ilo = _iv(m.group(1)) if m else lo
            m = re.search(r"<interiorHighIndex>([^<]*)<", p)
            ihi = _iv(m.group(1)) if m else hi
            m = re.search(r"<proc>([^<]*)<", p)
            patches.append((int(g("id", p)), lo, hi, ilo, ihi, int(m.group(1)) if m else -1))
        levels.append(dict(anchor=_vv(g("anchor")), dx=_vv(g("cellspacing")), extra=_iv(g("extraCells")), patches=patches))
    return levels


def timesteps(uda):
    x = open(os.path.join(uda, "index.xml")).read()
    return [(int(n), href) for href, n in re.findall(r'<timestep href="([^"]+)"[^>]*>(\d+)</timestep>', x)]


def read_patch_var(uda, tdir, level, pid, name, shape_zyx, matl=0, proc=-1):
    """the files are named after the writing rank: lL/p<proc>.xml lists every patch that rank owned"""
    ldir = os.path.join(uda, tdir, f"l{level}")
    x = open(os.path.join(ldir, f"p{(proc if proc >= 0 else pid):05d}.xml")).read()
    for t, body in re.findall(r'<Variable type="([^"]+)">(.*?)</Variable>', x, re.S):
        g = lambda tag: re.search(rf"<{tag}>([^<]*)</{tag}>", body)
        if g("variable").group(1) != name or int(g("index").group(1)) != matl or int(g("patch").group(1)) != pid:
            continue
        sub = re.search(r"&lt;(\w+)&gt;", t).group(1)
        start, end = int(g("start").group(1)), int(g("end").group(1))
        with open(os.path.join(ldir, g("filename").group(1)), "rb") as f:
            f.seek(start)
            raw = f.read(end - start)
        if g("compression") and g("compression").group(1) == "gzip":
            (n,) = struct.unpack("<Q", raw[:8])
            raw = zlib.decompress(raw[8:])
            assert len(raw) == n
        if sub == "Stencil7":
            return np.frombuffer(raw, np.float64).reshape(shape_zyx + (7,))
        return np.frombuffer(raw, _TYPES[sub]).reshape(shape_zyx)
    raise KeyError(name)


def read_level_var(uda, name, level=0, step=-1, matl=0):
    """assemble a variable over the whole level window [lo,hi) incl. extra cells; returns (array [z][y][x], lo, hi)"""
    n, href = timesteps(uda)[step]
    tdir = os.path.dirname(href)
    lev = read_grid(uda, tdir)[level]
    lo = tuple(min(p[1][d] for p in lev["patches"]) for d in range(3))
    hi = tuple(max(p[2][d] for p in lev["patches"]) for d in range(3))
    out = None
    for pid, plo, phi, _, _, proc in lev["patches"]:
        a = read_patch_var(uda, tdir, level, pid, name, (phi[2] - plo[2], phi[1] - plo[1], phi[0] - plo[0]), matl, proc)
        if out is None:
            out = np.zeros((hi[2] - lo[2], hi[1] - lo[1], hi[0] - lo[0]) + a.shape[3:], a.dtype)
        out[plo[2] - lo[2]:phi[2] - lo[2], plo[1] - lo[1]:phi[1] - lo[1], plo[0] - lo[0]:phi[0] - lo[0]] = a
    return out, lo, hi


def write_uda(uda, fields, lo, hi, interior_lo, interior_hi, patches, anchor, dx, extra=(1, 1, 1), timestep=1, time=0.0, gzip=True, nprocs=None):
    """fields: {name: array [z][y][x] over [lo,hi)} (int32 / float32 / float64); patches = (px,py,pz) split of the interior,
    extra cells attached to the patches on the domain faces (the layout sus writes).  nprocs: ranks that "wrote" the archive (default one
    per patch); patch p belongs to rank p % nprocs and lives in that rank's l0/p<rank>.{xml,data}"""
    tdir = f"t{timestep:05d}"
    os.makedirs(os.path.join(uda, tdir, "l0"), exist_ok=True)
    names = {np.dtype(np.int32): "int", np.dtype(np.float32): "float", np.dtype(np.float64): "double"}
    plist = []
    pid = 0
    edges = [[interior_lo[d] + ((interior_hi[d] - interior_lo[d]) * i) // patches[d] for i in range(patches[d] + 1)] for d in range(3)]
    for pi in range(patches[0]):            # x slowest: the order sus numbers patches in (restartUda_RT.uda)
        for pj in range(patches[1]):
            for pk in range(patches[2]):
                ilo = (edges[0][pi], edges[1][pj], edges[2][pk])
                ihi = (edges[0][pi + 1], edges[1][pj + 1], edges[2][pk + 1])
                idx = (pi, pj, pk)
                plo = tuple(ilo[d] - (extra[d] if idx[d] == 0 else 0) for d in range(3))
                phi = tuple(ihi[d] + (extra[d] if idx[d] == patches[d] - 1 else 0) for d in range(3))
                plist.append((pid, plo, phi, ilo, ihi))
                pid += 1
    nprocs = nprocs or len(plist)
    for rank in range(nprocs):
      xml, off = [], 0
      with open(os.path.join(uda, tdir, "l0", f"p{rank:05d}.data"), "wb") as f:
        for pid, plo, phi, ilo, ihi in plist:
            if pid % nprocs != rank:
                continue
            for name, arr in fields.items():
                sub = arr[plo[2] - lo[2]:phi[2] - lo[2], plo[1] - lo[1]:phi[1] - lo[1], plo[0] - lo[0]:phi[0] - lo[0]]
                raw = np.ascontiguousarray(sub).tobytes()
                blob = struct.pack("<Q", len(raw)) + zlib.compress(raw) if gzip else raw
                f.write(blob)
                comp = "    <compression>gzip</compression>\n" if gzip else ""
                xml.append(f'  <Variable type="CCVariable&lt;{names[arr.dtype]}&gt;">\n    <variable>{name}</variable>\n    <index>0</index>\n'
                           f"    <patch>{pid}</patch>\n    <start>{off}</start>\n{comp}    <end>{off + len(blob)}</end>\n"
                           f"    <filename>p{rank:05d}.data</filename>\n  </Variable>\n")
                off += len(blob)
      with open(os.path.join(uda, tdir, "l0", f"p{rank:05d}.xml"), "w") as f:
            f.write('<?xml version="1.0" encoding="UTF-8"?>\n<Uintah_Output>\n' + "".join(xml) + "</Uintah_Output>\n")
    v3 = lambda v: "[" + ", ".join(repr(float(x)) for x in v) + "]"
    i3 = lambda v: "[" + ", ".join(str(int(x)) for x in v) + "]"
    with open(os.path.join(uda, tdir, "timestep.xml"), "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<Uintah_timestep>\n  <Meta>\n    <endianness>little_endian</endianness>\n    <nBits>64</nBits>\n'
                f"    <numProcs>{nprocs}</numProcs>\n  </Meta>\n  <Time>\n    <timestepNumber>{timestep}</timestepNumber>\n"
                f"    <currentTime>{time!r}</currentTime>\n    <oldDelt>1e-06</oldDelt>\n  </Time>\n  <Grid>\n    <numLevels>1</numLevels>\n    <Level>\n"
                f"      <numPatches>{len(plist)}</numPatches>\n      <extraCells>{i3(extra)}</extraCells>\n      <anchor>{v3(anchor)}</anchor>\n      <id>0</id>\n"
                f"      <cellspacing>{v3(dx)}</cellspacing>\n")
        for pid, plo, phi, ilo, ihi in plist:
            f.write(f"      <Patch>\n        <id>{pid}</id>\n        <proc>{pid % nprocs}</proc>\n        <lowIndex>{i3(plo)}</lowIndex>\n        <highIndex>{i3(phi)}</highIndex>\n")
            if ilo != plo:
                f.write(f"        <interiorLowIndex>{i3(ilo)}</interiorLowIndex>\n")
            if ihi != phi:
                f.write(f"        <interiorHighIndex>{i3(ihi)}</interiorHighIndex>\n")
            f.write("      </Patch>\n")
        f.write("    </Level>\n  </Grid>\n  <Data>\n" + "".join(f'    <Datafile href="l0/p{r:05d}.xml" proc="{r}"/>\n' for r in range(nprocs)) +
                "  </Data>\n</Uintah_timestep>\n")
    with open(os.path.join(uda, "index.xml"), "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<Uintah_DataArchive>\n  <Meta>\n    <endianness>little_endian</endianness>\n    <nBits>64</nBits>\n  </Meta>\n  <variables>\n')
        for name, arr in fields.items():
            f.write(f'    <variable type="CCVariable&lt;{names[arr.dtype]}&gt;" name="{name}"/>\n')
        f.write(f'  </variables>\n  <timesteps>\n    <timestep href="{tdir}/timestep.xml" time="{time!r}" oldDelt="1e-06">{timestep}</timestep>\n  </timesteps>\n</Uintah_DataArchive>\n')
    return plist
