"""On-disk formats either side of the microstate path.

Everything here is host-side text handling; nothing touches the GPU.  It exists so that a
user without pDynamo/MEAD can feed the engine from the files the reference itself writes:

* MEAD solver logs ``site_<label>.out`` / ``model_<label>.out``
  (grammar: ``MEADOutputFileReader.py:36-58``; assembly into Gintr and W:
  ``InstanceMEAD.py:112-137``, ``Instance.py:100-108``, ``SiteMEAD.py:101-105``),
* site parameter files ``parameters/sites/*.yaml`` and ``*.est``
  (``TemplatesLibrary.py``, ``ESTFileReader.py:43-51``),
* GMCT-compatible tables written by ``CEModel.WriteGintr`` / ``CEModel.WriteW``
  (``CEModel.py:363-438``; the 8-decimal ``job.gint`` / ``job.inter`` fixtures),
* titration-curve files ``<site>_<instance>.dat`` (``TitrationCurves.py:169-194``).
"""

from __future__ import annotations

import os
import re
from dataclasses import dataclass, field

import numpy as np

from .constants import DEFAULT_TEMPERATURE


# ---------------------------------------------------------------------------------------
# Model record: the plain-array input of the hot path
# ---------------------------------------------------------------------------------------
@dataclass
class ModelRecord:
    """Everything the microstate path needs, as plain arrays.

    ``site_first[i]..site_last[i]`` is the contiguous, ascending range of *global* instance
    indices of site ``i`` (the reference's ``TitrSite.indexFirst/indexLast``,
    ``StateVector.h:26-35``).  ``interactions`` is the raw, possibly asymmetric
    instance-by-instance matrix (``EnergyModel.h:46``), row = the instance MEAD was run for.
    """

    name: str
    site_first: np.ndarray
    site_last: np.ndarray
    protons: np.ndarray
    gmodel: np.ndarray
    gintr: np.ndarray
    interactions: np.ndarray
    temperature: float = DEFAULT_TEMPERATURE
    site_labels: list = field(default_factory=list)   # (segName, resName, resSerial)
    inst_labels: list = field(default_factory=list)   # per global instance

    @property
    def nsites(self) -> int:
        return int(self.site_first.shape[0])

    @property
    def ninstances(self) -> int:
        return int(self.gintr.shape[0])

    @property
    def nstates(self) -> int:
        n = 1
        for a, b in zip(self.site_first, self.site_last):
            n *= int(b) - int(a) + 1
        return n

    def symmetric(self) -> np.ndarray:
        """(W + W^T)/2, what ``EnergyModel_SymmetrizeInteractions`` stores (EnergyModel.c:85-87)."""
        w = np.asarray(self.interactions, dtype=np.float64)
        return (w + w.T) * 0.5

    def validate(self) -> None:
        first, last = self.site_first, self.site_last
        if first.shape != last.shape or first.ndim != 1:
            raise ValueError("site_first/site_last must be 1-D and conformable")
        expect = 0
        for a, b in zip(first, last):
            if a != expect or b < a:
                raise ValueError("site instance ranges must be contiguous and ascending")
            expect = int(b) + 1
        n = self.ninstances
        if expect != n:
            raise ValueError("site ranges do not cover all instances")
        for arr in (self.protons, self.gmodel):
            if arr.shape != (n,):
                raise ValueError("per-instance arrays must have ninstances entries")
        if self.interactions.shape != (n, n):
            raise ValueError("interactions must be ninstances x ninstances")

    # -- persistence ------------------------------------------------------------------
    def to_npz_dict(self) -> dict:
        return dict(
            name=np.array(self.name),
            temperature=np.array(self.temperature, dtype=np.float64),
            site_first=self.site_first.astype(np.int32),
            site_last=self.site_last.astype(np.int32),
            protons=self.protons.astype(np.int32),
            gmodel=self.gmodel.astype(np.float64),
            gintr=self.gintr.astype(np.float64),
            interactions=self.interactions.astype(np.float64),
            site_labels=np.array(["%s|%s|%d" % tuple(s) for s in self.site_labels]),
            inst_labels=np.array(list(self.inst_labels)),
        )

    @classmethod
    def from_npz_dict(cls, d) -> "ModelRecord":
        labels = []
        for s in d["site_labels"].tolist():
            seg, res, serial = s.split("|")
            labels.append((seg, res, int(serial)))
        return cls(
            name=str(d["name"]),
            temperature=float(d["temperature"]),
            site_first=np.asarray(d["site_first"], dtype=np.int32),
            site_last=np.asarray(d["site_last"], dtype=np.int32),
            protons=np.asarray(d["protons"], dtype=np.int32),
            gmodel=np.asarray(d["gmodel"], dtype=np.float64),
            gintr=np.asarray(d["gintr"], dtype=np.float64),
            interactions=np.asarray(d["interactions"], dtype=np.float64),
            site_labels=labels,
            inst_labels=[str(x) for x in d["inst_labels"].tolist()],
        )


# ---------------------------------------------------------------------------------------
# MEAD solver logs
# ---------------------------------------------------------------------------------------
def parse_mead_output(text: str) -> dict:
    """Parse one ``my_2diel_solver`` / ``my_3diel_solver`` log.

    Follows ``MEADOutputFileReader.Parse`` (MEADOutputFileReader.py:23-65): the last token of the
    "Self energy of" line is the Born energy, the last token of "Interaction energy of" is the
    background energy, and after "Interaction energies of" (+2 skipped lines) every line
    ``site instance : energy`` up to "Total runtime" is one interaction.
    """
    out = {"interactions": []}
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        line = lines[i]
        if line.startswith("Self energy of"):
            out["born"] = float(line.split()[-1])
        elif line.startswith("Interaction energy of"):
            out["back"] = float(line.split()[-1])
        elif line.startswith("Interaction energies of"):
            i += 3
            while i < len(lines) and not lines[i].startswith("Total runtime"):
                tokens = lines[i].split()
                if len(tokens) >= 4:
                    out["interactions"].append((int(tokens[0]), int(tokens[1]), float(tokens[3])))
                i += 1
        i += 1
    return out


def _parse_simple_site_yaml(text: str) -> dict:
    """Minimal reader for ``parameters/sites/*.yaml`` (label / Gmodel / protons per instance)."""
    import yaml

    doc = yaml.safe_load(text)
    instances = [(str(inst["label"]), float(inst["Gmodel"]), int(inst["protons"])) for inst in doc["instances"]]
    return {str(doc["site"]): instances}


def parse_est(text: str, site_label: str) -> dict:
    """EST parameter file: rows ``label``, ``Gmodel``, ``proton`` (ESTFileReader.py:43-51)."""
    labels, gmodels, protons = None, None, None
    for line in text.splitlines():
        tokens = line.split()
        if not tokens:
            continue
        if tokens[0] == "Gmodel":
            gmodels = [float(t) for t in tokens[1:]]
        elif tokens[0] == "proton":
            protons = [int(t) for t in tokens[1:]]
        elif tokens[0] == "label":
            labels = tokens[1:]
    return {site_label: list(zip(labels, gmodels, protons))}


def load_site_parameters(yaml_directory: str, est_files=()) -> dict:
    """``{resName: [(label, Gmodel, protons), ...]}`` from a ``parameters/sites`` tree plus EST files."""
    table = {}
    for name in sorted(os.listdir(yaml_directory)):
        if name.endswith(".yaml"):
            with open(os.path.join(yaml_directory, name)) as handle:
                table.update(_parse_simple_site_yaml(handle.read()))
    for path in est_files:
        label = os.path.splitext(os.path.basename(path))[0]
        with open(path) as handle:
            table.update(parse_est(handle.read(), label))
    return table


_INSTANCE_ROW = re.compile(
    r"^\s*(\S+)\s+(\S+)\s+(-?\d+)\s+(\S+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s*$"
)


def parse_instance_table(log_text: str) -> list:
    """Rows of the "Instance of a site ... Gintr" table of a reference log (``Instance._TableEntry``,
    Instance.py:167-183): ``(seg, res, serial, label, Gborn_m, Gback_m, Gborn_p, Gback_p, Gmodel, Gintr)``.

    Only the first table of the log is returned (lysozyme_compare logs hold two models).
    """
    rows = []
    inside = False
    for line in log_text.splitlines():
        if "Instance of a site" in line and "Gintr" in line:
            if rows:
                break
            inside = True
            continue
        if inside:
            match = _INSTANCE_ROW.match(line)
            if match:
                g = match.groups()
                rows.append((g[0], g[1], int(g[2]), g[3]) + tuple(float(x) for x in g[4:]))
            elif rows and line.strip() and not line.startswith("-"):
                break
    return rows


def model_from_mead_directory(name, log_text, mead_directory, parameters, temperature=DEFAULT_TEMPERATURE) -> ModelRecord:
    """Rebuild ``(Gintr, protons, W)`` from cached MEAD logs, the way the reference does on a rerun.

    Site / instance order comes from the log's instance table; energies come from
    ``<mead>/<seg>/<RES><serial>/{model,site}_<label>.out``:

    * ``Gmodel = template.Gmodel * T/300``                               (SiteMEAD.py:101)
    * ``Gintr  = Gmodel + (Gborn_p - Gborn_m) + (Gback_p - Gback_m)``      (Instance.py:108)
    * ``W[row][col]`` in (site, instance) order, same-site entries zeroed  (InstanceMEAD.py:112-137)
    """
    rows = parse_instance_table(log_text)
    if not rows:
        raise ValueError("no instance table found in log")
    site_labels, site_first, site_last, inst_labels, inst_site = [], [], [], [], []
    for idx, (seg, res, serial, label, *_rest) in enumerate(rows):
        key = (seg, res, serial)
        if not site_labels or site_labels[-1] != key:
            site_labels.append(key)
            site_first.append(idx)
            site_last.append(idx)
        else:
            site_last[-1] = idx
        inst_labels.append(label)
        inst_site.append(len(site_labels) - 1)
    ninst = len(rows)
    protons = np.zeros(ninst, dtype=np.int32)
    gmodel = np.zeros(ninst, dtype=np.float64)
    gintr = np.zeros(ninst, dtype=np.float64)
    w = np.zeros((ninst, ninst), dtype=np.float64)
    for idx, (seg, res, serial, label, *_rest) in enumerate(rows):
        template = {lab: (gm, pr) for lab, gm, pr in parameters[res]}
        gm, pr = template[label]
        gmodel[idx] = gm * temperature / 300.0
        protons[idx] = pr
        directory = os.path.join(mead_directory, seg, "%s%d" % (res, serial))
        with open(os.path.join(directory, "model_%s.out" % label)) as handle:
            model_log = parse_mead_output(handle.read())
        with open(os.path.join(directory, "site_%s.out" % label)) as handle:
            site_log = parse_mead_output(handle.read())
        gintr[idx] = gmodel[idx] + (site_log["born"] - model_log["born"]) + (site_log["back"] - model_log["back"])
        if len(site_log["interactions"]) != ninst:
            raise ValueError("%s: expected %d interactions, found %d" % (directory, ninst, len(site_log["interactions"])))
        for col, (site_index, _inst_index, energy) in enumerate(site_log["interactions"]):
            w[idx, col] = 0.0 if site_index == inst_site[idx] else energy
    record = ModelRecord(
        name=name,
        site_first=np.array(site_first, dtype=np.int32),
        site_last=np.array(site_last, dtype=np.int32),
        protons=protons,
        gmodel=gmodel,
        gintr=gintr,
        interactions=w,
        temperature=temperature,
        site_labels=site_labels,
        inst_labels=inst_labels,
    )
    record.validate()
    return record


# ---------------------------------------------------------------------------------------
# GMCT-compatible tables (CEModel.WriteGintr / CEModel.WriteW)
# ---------------------------------------------------------------------------------------
def read_gintr(path: str):
    """Rows ``(siteID, instID, siteLabel, instLabel, Gintr, protons)``; keyed on column position
    because the header names changed between reference revisions (fixtures vs CEModel.py:411-417)."""
    rows = []
    with open(path) as handle:
        for line in handle:
            if line.startswith("#") or not line.strip():
                continue
            t = line.split()
            rows.append((int(t[0]), int(t[1]), t[2], t[3], float(t[4]), int(t[5])))
    return rows


def read_w(path: str):
    """Rows ``(siteA, instA, labSiteA, labInstA, siteB, instB, labSiteB, labInstB, Wsym, W, Werr)``
    (CEModel.py:366-377)."""
    rows = []
    with open(path) as handle:
        for line in handle:
            if line.startswith("#") or not line.strip():
                continue
            t = line.split()
            rows.append((int(t[0]), int(t[1]), t[2], t[3], int(t[4]), int(t[5]), t[6], t[7], float(t[8]), float(t[9]), float(t[10])))
    return rows


def model_from_gmct_tables(name, gint_path, inter_path, temperature=DEFAULT_TEMPERATURE) -> ModelRecord:
    """Build a model record from ``job.gint`` + ``job.inter`` (raw ``Wij`` column, so that
    symmetrisation is re-done by the engine exactly as for MEAD input)."""
    grows = read_gintr(gint_path)
    index = {}
    site_labels, site_first, site_last, inst_labels = [], [], [], []
    for k, (sid, iid, slabel, ilabel, _g, _p) in enumerate(grows):
        index[(sid, iid)] = k
        if sid - 1 == len(site_labels):
            parts = slabel.split("_")
            match = re.match(r"([A-Za-z]+)(-?\d+)$", parts[-1])
            seg = "_".join(parts[:-1])
            site_labels.append((seg, match.group(1), int(match.group(2))) if match else (seg, parts[-1], 0))
            site_first.append(k)
            site_last.append(k)
        else:
            site_last[-1] = k
        inst_labels.append(ilabel)
    ninst = len(grows)
    w = np.zeros((ninst, ninst), dtype=np.float64)
    for (sa, ia, _1, _2, sb, ib, _3, _4, _wsym, wraw, _werr) in read_w(inter_path):
        w[index[(sa, ia)], index[(sb, ib)]] = wraw
    record = ModelRecord(
        name=name,
        site_first=np.array(site_first, dtype=np.int32),
        site_last=np.array(site_last, dtype=np.int32),
        protons=np.array([r[5] for r in grows], dtype=np.int32),
        gmodel=np.zeros(ninst, dtype=np.float64),
        gintr=np.array([r[4] for r in grows], dtype=np.float64),
        interactions=w,
        temperature=temperature,
        site_labels=site_labels,
        inst_labels=inst_labels,
    )
    record.validate()
    return record


def _table_format(items):
    header = "# "
    form = ""
    for label, width, digits in items:
        header = "%s%s" % (header, label.center(width))
        if digits < 0:
            form = "%s%%%ds" % (form, width)
        elif digits < 1:
            form = "%s%%%dd" % (form, width)
        else:
            form = "%s%%%d.%df" % (form, width, digits)
    return header + "\n", form + "\n"


def format_gintr_lines(site_names, inst_rows, precision=4):
    """Lines of a ``gintr.dat`` (CEModel.WriteGintr, CEModel.py:405-438).  ``inst_rows`` are
    ``(siteIndex, instIndex, instLabel, Gintr, protons)``.  The reference ignores its own
    ``precision`` argument and always prints 4 decimals; here the argument is honoured, with the
    reference's 4 as the default."""
    items = (("siteID", 8, 0), ("instID", 8, 0), ("siteLabel", 14, -1), ("instLabel", 10, -1),
             ("Gintr", max(12, precision + 8), precision), ("protons", 8, 0))
    header, form = _table_format(items)
    lines = [header]
    for site_index, inst_index, inst_label, gintr, protons in inst_rows:
        lines.append(form % (site_index + 1, inst_index + 1, site_names[site_index], inst_label, gintr, protons))
    return lines


def format_w_lines(site_names, inst_rows, wsym, wraw, precision=4):
    """Lines of a ``W.dat`` (CEModel.WriteW, CEModel.py:363-402).  ``inst_rows`` are
    ``(siteIndex, instIndex, instLabel)`` in global-instance order."""
    items = (("idSiteA", 8, 0), ("idInstA", 8, 0), ("labSiteA", 14, -1), ("labInstA", 10, -1),
             ("idSiteB", 8, 0), ("idInstB", 8, 0), ("labSiteB", 14, -1), ("labInstB", 10, -1),
             ("Wij_symm", max(10, precision + 8), precision), ("Wij", max(10, precision + 8), precision),
             ("Wij_err", max(10, precision + 8), precision))
    header, form = _table_format(items)
    lines = [header]
    for a, (sa, ia, la) in enumerate(inst_rows):
        for b, (sb, ib, lb) in enumerate(inst_rows):
            deviation = (wraw[a, b] + wraw[b, a]) * 0.5 - wraw[a, b]   # EnergyModel_GetDeviation, EnergyModel.c:169-176
            lines.append(form % (sa + 1, ia + 1, site_names[sa], la, sb + 1, ib + 1, site_names[sb], lb,
                                 wsym[a, b], wraw[a, b], deviation))
    return lines


# ---------------------------------------------------------------------------------------
# Curve files
# ---------------------------------------------------------------------------------------
def read_curve_file(path: str) -> np.ndarray:
    """``%f %f`` lines of one ``<site>_<instance>.dat`` (TitrationCurves.py:188-191) -> (n, 2) array."""
    return np.loadtxt(path, dtype=np.float64, ndmin=2)


def read_curves_directory(directory: str, record: ModelRecord):
    """All curve files of a model, ordered by global instance -> ``(pH[n], p[n, ninst])``."""
    columns = []
    ph = None
    for site_index, (seg, res, serial) in enumerate(record.site_labels):
        for g in range(int(record.site_first[site_index]), int(record.site_last[site_index]) + 1):
            name = "%s_%s%d_%s.dat" % (seg, res, serial, record.inst_labels[g])
            data = read_curve_file(os.path.join(directory, name))
            if ph is None:
                ph = data[:, 0]
            columns.append(data[:, 1])
    return ph, np.stack(columns, axis=1)
