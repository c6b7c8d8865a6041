"""AutoBench's output format: Stats.xlsx with one worksheet per scene and the 18 columns of
AutoBench/benchruntable.h:28-49 / benchruntable.cpp:87-145, written by the dependency-free writer in
flipsolver2d_b200/host/benchruntable.h. Read back with Python's zipfile + XML parser (what any spreadsheet does)."""
import os
import sys
import xml.etree.ElementTree as ET
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flipsolver2d_b200 import host_api  # noqa: E402

NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main",
      "r": "http://schemas.openxmlformats.org/officeDocument/2006/relationships"}
HEADERS = ["Step number", "Substeps", "Frame time", "Advection", "Decomposition", "Density correction", "Particle rebin",
           "Particle to grid", "Grid update", "After transfer", "Pressure", "Viscosity", "After-visc pressure",
           "Particle update", "Particle reseeding", "Pressure iterations", "Density iterations", "Viscosity iterations"]


def _sheet_rows(z, path):
    root = ET.fromstring(z.read(path))
    rows = []
    for row in root.find("m:sheetData", NS).findall("m:row", NS):
        cells = []
        for c in row.findall("m:c", NS):
            if c.get("t") == "inlineStr":
                cells.append(c.find("m:is", NS).find("m:t", NS).text)
            else:
                cells.append(float(c.find("m:v", NS).text))
        rows.append(cells)
    return rows


def test_stats_xlsx_layout_and_values(tmp_path):
    rng = np.random.default_rng(5)
    scenes = {}
    for name, frames in (("dam_break", 7), ("smoke_test", 3), ("a<b&c", 1)):
        t = rng.random((frames, 18)) * 100.0
        t[:, 0] = np.arange(1, frames + 1)          # step number: 1-based (benchruntable.cpp:66)
        t[:, 1] = rng.integers(1, 11, frames)       # substeps
        t[:, 15:] = rng.integers(0, 201, (frames, 3))
        scenes[name] = t
    path = tmp_path / "Stats.xlsx"
    host_api.write_stats_xlsx(path, scenes)
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None                                   # CRCs of all stored members are right
        names = z.namelist()
        assert "[Content_Types].xml" in names and "xl/workbook.xml" in names and "_rels/.rels" in names
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        sheets = [(s.get("name"), s.get("{%s}id" % NS["r"])) for s in wb.find("m:sheets", NS).findall("m:sheet", NS)]
        rels = {r.get("Id"): r.get("Target") for r in ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))}
        # OpenXLSX's default empty "Sheet1" first, then the scenes in std::map (sorted) order
        assert [s[0] for s in sheets] == ["Sheet1"] + sorted(scenes.keys())
        assert _sheet_rows(z, "xl/" + rels[sheets[0][1]]) == []
        for name, rid in sheets[1:]:
            rows = _sheet_rows(z, "xl/" + rels[rid])
            assert rows[0] == HEADERS
            got = np.array(rows[1:], np.float64)
            assert got.shape == scenes[name].shape
            assert np.allclose(got, scenes[name], rtol=1e-8, atol=0)
        types = z.read("[Content_Types].xml").decode()
        assert types.count("worksheet+xml") == len(sheets)


def test_stats_xlsx_long_and_clashing_scene_names(tmp_path):
    long_a = "x" * 40 + "_one"
    long_b = "x" * 40 + "_two"   # same first 31 characters: worksheet names are capped at 31 and must stay unique
    scenes = {long_a: np.ones((1, 18)), long_b: 2 * np.ones((1, 18)), "with/slash": 3 * np.ones((1, 18))}
    path = tmp_path / "Stats.xlsx"
    host_api.write_stats_xlsx(path, scenes)
    with zipfile.ZipFile(path) as z:
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        names = [s.get("name") for s in wb.find("m:sheets", NS).findall("m:sheet", NS)]
        assert len(set(names)) == len(names) == 4
        assert all(len(n) <= 31 and not any(ch in n for ch in "[]:*?/\\") for n in names)
