#!/usr/bin/env python3
"""Generates the data fixtures that come from OpenDXMC's own data files.

Run in the build container (needs /root/reference); the outputs are committed because
/root/reference does not exist on the GPU box:

  opendxmc_b200/data/bowtiefilters.json   compact copy of R:data/bowtiefilters/bowtiefilters.json
                                          (41 filters, same "filters"/"filterdata"/"name" schema that
                                          R:src/libopendxmc/bowtiefilterreader.cpp:51-97 parses)
  opendxmc_b200/data/icrp_tables.json     organ -> (medium, density) and medium -> composition tables of the
                                          ICRP 110 / 143 phantoms, parsed with the rules of
                                          R:src/libopendxmc/icrpphantomimportpipeline.cpp:60-200
                                          (organ line: id, 50-char name field, medium id, density; media
                                          line: id, name, 13 mass-% columns for Z = 1,6,7,8,11,12,15,16,17,19,20,26,53)
  opendxmc_b200/data/icrp_shapes.json     phantom dimensions / spacings from R:src/app/icrpphantomimportwidget.cpp:65-140

These are INPUT fixtures only; the reference holds no expected outputs for the transport path.
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "opendxmc_b200", "data")
ZS = [1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 53]


def parse_organ_line(line):
    if not line[:1].isdigit():
        return None
    m = re.match(r"(\d+)", line)
    oid = int(m.group(1))
    if oid > 255:
        return None
    rest_start = m.end()
    pos = rest_start + 50
    name_end = None
    medium = None
    while pos < len(line):
        mm = re.match(r"\d+", line[pos:])
        if mm:
            medium = int(mm.group(0))
            name_end = pos
            pos += mm.end()
            break
        pos += 1
    if medium is None:
        return None
    dens = None
    while pos < len(line):
        mm = re.match(r"[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?", line[pos:])
        if mm:
            dens = float(mm.group(0))
            break
        pos += 1
    if not dens:
        return None
    name = line[rest_start:name_end].strip()
    if not name:
        return None
    return {"id": oid, "name": name, "medium": medium, "density": dens}


def parse_media_line(line):
    if not line[:1].isdigit():
        return None
    m = re.match(r"(\d+)", line)
    mid = int(m.group(1))
    pos = m.end()
    start = pos
    while pos < len(line) and not line[pos].isdigit():
        pos += 1
    if pos >= len(line):
        return None
    name = line[start:pos].strip()
    vals = line[pos:].split()
    if len(vals) < 13:
        return None
    try:
        w = [float(v) for v in vals[:13]]
    except ValueError:
        return None
    return {"id": mid, "name": name, "composition": {str(z): v for z, v in zip(ZS, w)}}


def main():
    os.makedirs(OUT, exist_ok=True)
    src = json.load(open(os.path.join(REF, "data/bowtiefilters/bowtiefilters.json")))
    filters = []
    for f in src["filters"]:
        data = [{"angle": d["angle"], "weight": d["weight"]} for d in f.get("filterdata", []) if "angle" in d and "weight" in d]
        if data and f.get("name"):
            filters.append({"name": f["name"], "filterdata": data})
    json.dump({"filters": filters}, open(os.path.join(OUT, "bowtiefilters.json"), "w"), separators=(",", ":"))
    print("bowtie filters:", len(filters))

    tables = {}
    base = os.path.join(REF, "data/phantoms/icrp")
    for ph in sorted(os.listdir(base)):
        d = os.path.join(base, ph)
        organs = [o for o in (parse_organ_line(l.rstrip("\n")) for l in open(os.path.join(d, f"{ph}_organs.dat"), errors="replace")) if o]
        media = [m for m in (parse_media_line(l.rstrip("\n")) for l in open(os.path.join(d, f"{ph}_media.dat"), errors="replace")) if m]
        tables[ph] = {"organs": organs, "media": media}
        print(ph, "organs", len(organs), "media", len(media), "distinct media used", len({o['medium'] for o in organs}))
    json.dump(tables, open(os.path.join(OUT, "icrp_tables.json"), "w"), separators=(",", ":"))

    # R:src/app/icrpphantomimportwidget.cpp:65-140  (spacing in mm, dimensions, folder)
    txt = open(os.path.join(REF, "src/app/icrpphantomimportwidget.cpp")).read()
    shapes = {}
    pat = re.compile(r"\.spacing = \{\s*([\d.]+),\s*([\d.]+),\s*([\d.]+)\s*\},\s*\.dimensions = \{\s*(\d+),\s*(\d+),\s*(\d+)\s*\},"
                     r"\s*\.name = \"([^\"]+)\",\s*\.filePrefix = QStringLiteral\(\"[^\"]+\"\),\s*\.folderPath = QStringLiteral\(\"([^\"]+)\"\)")
    for m in pat.finditer(txt):
        shapes[m.group(8)] = {"spacing_mm": [float(m.group(i)) for i in (1, 2, 3)], "dimensions": [int(m.group(i)) for i in (4, 5, 6)],
                              "name": m.group(7)}
    json.dump(shapes, open(os.path.join(OUT, "icrp_shapes.json"), "w"), separators=(",", ":"))
    print("phantom shapes:", len(shapes))


if __name__ == "__main__":
    main()
