"""Velodyne calibration db.xml (boost-serialization XML) writer / reader.

Schema as the reference reads it (HDLParser.cxx:771-858): `boost_serialization.DB.enabled_`
holds one <item> per laser (1 = enabled; their count selects the 16/32/64-laser behaviour) and
`boost_serialization.DB.points_` one <item><px> per laser with id_, rotCorrection_,
vertCorrection_, distCorrection_ (cm), vertOffsetCorrection_ (cm), horizOffsetCorrection_ (cm).
"""
from __future__ import annotations

import re

import numpy as np

from .synth import Calibration


def write_db_xml(path, calib):
    n = calib.n_rows
    out = ['<?xml version="1.0" encoding="UTF-8" standalone="yes" ?>',
           "<!DOCTYPE boost_serialization>",
           '<boost_serialization signature="serialization::archive" version="4">',
           '<DB class_id="0" tracking_level="1" version="0" object_id="_0">',
           "\t<distLSB_>0.2</distLSB_>",
           '\t<enabled_ class_id="3" tracking_level="0" version="0">',
           f"\t\t<count>{max(n, calib.n_enabled)}</count>", "\t\t<item_version>0</item_version>"]
    m = max(n, calib.n_enabled)
    for i in range(m):
        out.append(f"\t\t<item>{1 if i < calib.n_enabled else 0}</item>")
    out += ["\t</enabled_>", '\t<points_ class_id="5" tracking_level="0" version="0">',
            f"\t\t<count>{n}</count>", "\t\t<item_version>1</item_version>"]
    for i in range(n):
        out += ['\t\t<item class_id="6" tracking_level="0" version="1">',
                '\t\t\t<px class_id="7" tracking_level="1" version="1" object_id="_%d">' % (i + 1),
                f"\t\t\t\t<id_>{i}</id_>",
                f"\t\t\t\t<rotCorrection_>{float(calib.rot_deg[i])!r}</rotCorrection_>",
                f"\t\t\t\t<vertCorrection_>{float(calib.vert_deg[i])!r}</vertCorrection_>",
                f"\t\t\t\t<distCorrection_>{float(calib.dist_cm[i])!r}</distCorrection_>",
                f"\t\t\t\t<vertOffsetCorrection_>{float(calib.voff_cm[i])!r}</vertOffsetCorrection_>",
                f"\t\t\t\t<horizOffsetCorrection_>{float(calib.hoff_cm[i])!r}</horizOffsetCorrection_>",
                "\t\t\t</px>", "\t\t</item>"]
    out += ["\t</points_>", "</DB>", "</boost_serialization>", ""]
    text = "\n".join(out)
    with open(path, "w") as f:
        f.write(text)


def read_db_xml(path):
    """Calibration from a db.xml (values parsed the way atof would)."""
    text = open(path).read()
    en = re.search(r"<enabled_[^>]*>(.*?)</enabled_>", text, re.S)
    n_enabled = sum(1 for v in re.findall(r"<item>\s*([-\d]+)\s*</item>", en.group(1))
                    if int(v) == 1) if en else 0
    rows = {}
    for px in re.findall(r"<px[^>]*>(.*?)</px>", text, re.S):
        def field(name, default=0.0):
            m = re.search(rf"<{name}>\s*([^<\s]+)\s*</{name}>", px)
            return float(m.group(1)) if m else default
        idx = int(field("id_", -1))
        if idx >= 0:
            rows[idx] = (field("rotCorrection_"), field("vertCorrection_"),
                         field("distCorrection_"), field("vertOffsetCorrection_"),
                         field("horizOffsetCorrection_"))
    n = max(rows) + 1 if rows else 0
    arr = np.zeros((5, n))
    for i, r in rows.items():
        arr[:, i] = r
    return Calibration(arr[0], arr[1], arr[2], arr[3], arr[4], n_enabled)
