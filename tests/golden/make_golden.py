#!/usr/bin/env python
"""Extract the reference's own fixtures into small, committed golden files.

Run in the build container only (it reads ``/root/reference``, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every fixture ``tests/<name>/results.tgz`` of the reference it writes ``tests/golden/<name>.npz`` with

* the model record rebuilt from the cached MEAD logs (site ranges, protons, Gmodel, Gintr, raw W), and
* the golden outputs the reference's logs/curve files hold for the microstate path: ``Gmicro`` lists,
  analytic and MC titration curves, pH-7 probability tables, double-move pair lists, and (where
  shipped) the 8-decimal ``job.gint`` / ``job.inter`` tables.

Nothing here is executed by the product; the files are test data.
"""

import io
import os
import re
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from pcetk_b200 import formats  # noqa: E402

REFERENCE = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# name -> (log file inside results/, analytic-curves dir, MC-curves dirs)
# NB twosites: directory names are swapped relative to their content (twosites.py:45-53 attaches
# the MC sampler before writing "curves_analytic"); see SURVEY.md section 0.
FIXTURES = {
    "twosites": dict(log="twosites.out", analytic="curves_mc", mc=["curves_analytic"]),
    "lysozyme": dict(log="lysozyme.out", analytic="curves_analytic", mc=["curves_mc"]),
    "defensin": dict(log="defensin.out", analytic="curves_analytic", mc=["curves_custom", "curves_gmct"]),
    "lysozyme_full_test": dict(log="lysozyme.out", analytic="curves_analytic", mc=["curves_custom", "curves_gmct"]),
    "rubredoxin": dict(log="rubredoxin.out", analytic="curves_analytic", mc=["curves_custom", "curves_gmct"]),
    "ifp20": dict(log="ifp20.out", analytic=None, mc=["curves"]),
}


def find(root, name):
    for base, _dirs, files in os.walk(root):
        if name in files:
            return os.path.join(base, name)
    return None


def parse_gmicro(log_text):
    """All 'Gmicro = x' titles (twosites.py:24-29 style) in order of appearance."""
    return [float(x) for x in re.findall(r"Gmicro\s*=\s*(-?\d+\.\d+)", log_text)]


def parse_state_gmicro_table(log_text):
    """The 'State  Gmicro' two-column table (first N microstate energies; defensin.py, lysozyme_full_test)."""
    values = []
    lines = log_text.splitlines()
    for i, line in enumerate(lines):
        if re.match(r"^\s*State\s+Gmicro\s*$", line):
            for row in lines[i + 2:]:
                m = re.match(r"^\s*(\d+)\s+(-?\d+\.\d+)\s*$", row)
                if not m:
                    break
                values.append(float(m.group(2)))
            break
    return values


def parse_pairs(log_text):
    """First 'Found N pair(s) of strongly interacting sites' block (MCModelDefault.pyx:176-194)."""
    lines = log_text.splitlines()
    for i, line in enumerate(lines):
        m = re.match(r"^Found (\d+) pairs? of strongly interacting sites", line)
        if m:
            pairs = []
            for row in lines[i + 1:i + 1 + int(m.group(1))]:
                pm = re.match(r"^\s*(\S+)\s+(\S+)\s+(-?\d+)\s+--\s+(\S+)\s+(\S+)\s+(-?\d+)\s+:\s+(-?\d+\.\d+)", row)
                pairs.append(((pm.group(1), pm.group(2), int(pm.group(3))), (pm.group(4), pm.group(5), int(pm.group(6))), float(pm.group(7))))
            return pairs
    return []


def parse_probability_tables(log_text, record):
    """Every 'Probabilities of instances' table (CEModel.SummaryProbabilities, CEModel.py:220-274) -> list of p[ninst] (4 dp)."""
    tables = []
    lines = log_text.splitlines()
    i = 0
    site_index = {s: k for k, s in enumerate(record.site_labels)}
    while i < len(lines):
        if "Probabilities of instances" in lines[i]:
            p = np.full(record.ninstances, np.nan)
            j = i + 2
            while j < len(lines) and not lines[j].startswith("---"):
                t = lines[j].split()
                key = (t[0], t[1], int(t[2]))
                k = site_index[key]
                first = int(record.site_first[k])
                values = t[3:]
                for q in range(len(values) // 2):
                    p[first + q] = float(values[2 * q + 1])
                j += 1
            tables.append(p)
            i = j
        i += 1
    return tables


def parse_substate_table(log_text):
    """First 'State Gmicro Charge Protons <sites...>' table (Substate.Summary, Substate.py:122-168)."""
    lines = log_text.splitlines()
    for i, line in enumerate(lines):
        if re.match(r"^\s*State\s+Gmicro\s+Charge\s+Protons", line):
            sites = re.findall(r"(\S+) (\S+) (\d+)", line[line.index("Protons") + 7:])
            energies, protons, labels = [], [], []
            for row in lines[i + 2:]:
                t = row.split()
                if len(t) < 4 + len(sites) or not t[0].isdigit():
                    break
                energies.append(float(t[1])); protons.append(int(t[3])); labels.append("|".join(t[4:4 + len(sites)]))
            return [(a, b, int(c)) for a, b, c in sites], energies, protons, labels
    return None


def parse_halfpk_table(log_text, record):
    """First 'pK1/2 values of instances' table (TitrationCurves.PrintHalfpKs): first pK1/2 per instance, NaN for n/a."""
    lines = log_text.splitlines()
    site_index = {s: k for k, s in enumerate(record.site_labels)}
    for i, line in enumerate(lines):
        if "pK1/2 values of instances" in line:
            out = np.full(record.ninstances, np.nan)
            for row in lines[i + 2:]:
                t = row.split()
                if len(t) < 5 or row.startswith("---"):
                    break
                key = (t[0], t[1], int(t[2]))
                k = site_index[key]
                labels = record.inst_labels[int(record.site_first[k]):int(record.site_last[k]) + 1]
                rest, g = t[3:], int(record.site_first[k])
                pos = 0
                for lab in labels:
                    assert rest[pos] == lab, (rest, lab)
                    pos += 1
                    if pos < len(rest) and rest[pos] != "n/a" and rest[pos] not in labels:
                        out[g] = float(rest[pos])
                    while pos < len(rest) and rest[pos] not in labels[labels.index(lab) + 1:labels.index(lab) + 2]:
                        pos += 1
                    g += 1
            return out
    return None


def parse_comparison_table(log_text, record, column):
    """The 'Comparison between previous and present results' table of lysozyme_compare (lysozyme.py:62-82 prints the first
    pK1/2 of every instance for both structures, 1 decimal, 'n/a' where a curve never crosses 0.5) -> pK1/2 per instance of
    ``column`` (0 = pKold, 1 = pKnew), NaN for n/a."""
    lines = log_text.splitlines()
    start = [i for i, line in enumerate(lines) if "Comparison between previous and present results" in line][0]
    out = np.full(record.ninstances, np.nan)
    site_index = {(res, serial): k for k, (_seg, res, serial) in enumerate(record.site_labels)}
    for row in lines[start + 4:]:
        if row.startswith("---"):
            break
        t = row.split()
        k = site_index[(t[0], int(t[1]))]
        ninst = int(record.site_last[k]) - int(record.site_first[k]) + 1
        pairs = t[5:]                                    # label value label value ... (old), then the same for new
        assert len(pairs) == 4 * ninst, row
        chosen = pairs[2 * ninst * column:2 * ninst * (column + 1)]
        for q in range(ninst):
            label, value = chosen[2 * q], chosen[2 * q + 1]
            assert label == record.inst_labels[int(record.site_first[k]) + q], (row, label)
            if value != "n/a":
                out[int(record.site_first[k]) + q] = float(value)
    return out


def extract_lysozyme_compare(parameters):
    """tests/lysozyme_compare: two lysozyme structures (7LYZ of 1977, 2LZT of 1990) in one run -- two models, two sets of
    MCModelDefault curves, and a table of the first pK1/2 of every instance for both."""
    name = "lysozyme_compare"
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(REFERENCE, "tests", name, "results.tgz")) as tar:
            tar.extractall(tmp)
        results = os.path.join(tmp, "results")
        with open(os.path.join(results, "lysozyme.out")) as handle:
            log_text = handle.read()
        marks = [m.start() for m in re.finditer(r"\*\*\* Now calculating:", log_text)]
        parts = [log_text[marks[0]:marks[1]], log_text[marks[1]:]]
        for column, (label, part) in enumerate(zip(("7LYZ", "2LZT"), parts)):
            record = formats.model_from_mead_directory("%s_%s" % (name, label), part, os.path.join(results, "mead", label), parameters)
            payload = record.to_npz_dict()
            payload["golden_gintr_4dp"] = np.array([r[-1] for r in formats.parse_instance_table(part)])
            pairs = parse_pairs(part)
            site_index = {s: k for k, s in enumerate(record.site_labels)}
            payload["golden_pairs"] = np.array([[site_index[a], site_index[b]] for a, b, _w in pairs], dtype=np.int32).reshape(-1, 2)
            payload["golden_pairs_wmax"] = np.array([w for _a, _b, w in pairs], dtype=np.float64)
            ph, p = formats.read_curves_directory(os.path.join(results, "curves", label), record)
            payload["golden_curve_ph"] = ph
            payload["golden_curve_mc_curves"] = p
            payload["golden_halfpk"] = parse_comparison_table(log_text, record, column)
            path = os.path.join(OUT, "%s_%s.npz" % (name, label))
            np.savez_compressed(path, **payload)
            print("%-20s nsites=%3d ninst=%3d nstates=%.4g pairs=%d halfpk=%d -> %s (%d KB)" % (
                name + "/" + label, record.nsites, record.ninstances, float(record.nstates), len(pairs),
                int(np.isfinite(payload["golden_halfpk"]).sum()), os.path.basename(path), os.path.getsize(path) // 1024))


def main():
    parameters_dir = os.path.join(REFERENCE, "parameters", "sites")
    est = [os.path.join(REFERENCE, "tests", "ifp20", n) for n in ("BLF.est", "ACB.est", "ACC.est")]
    parameters = formats.load_site_parameters(parameters_dir, est)
    for name, spec in FIXTURES.items():
        with tempfile.TemporaryDirectory() as tmp:
            with tarfile.open(os.path.join(REFERENCE, "tests", name, "results.tgz")) as tar:
                tar.extractall(tmp)
            results = os.path.join(tmp, "results")
            with open(os.path.join(results, spec["log"])) as handle:
                log_text = handle.read()
            record = formats.model_from_mead_directory(name, log_text, os.path.join(results, "mead"), parameters)
            payload = record.to_npz_dict()

            # Gintr as printed in the log (4 dp) is a cross-check of the rebuild
            rows = formats.parse_instance_table(log_text)
            payload["golden_gintr_4dp"] = np.array([r[-1] for r in rows])

            gm = parse_gmicro(log_text)
            if not gm:
                gm = parse_state_gmicro_table(log_text)
            payload["golden_gmicro"] = np.array(gm, dtype=np.float64)

            pairs = parse_pairs(log_text)
            site_index = {s: k for k, s in enumerate(record.site_labels)}
            payload["golden_pairs"] = np.array([[site_index[a], site_index[b]] for a, b, _w in pairs], dtype=np.int32).reshape(-1, 2)
            payload["golden_pairs_wmax"] = np.array([w for _a, _b, w in pairs], dtype=np.float64)

            tables = parse_probability_tables(log_text, record)
            if tables:
                payload["golden_prob_tables_ph7"] = np.stack(tables)

            if spec["analytic"]:
                ph, p = formats.read_curves_directory(os.path.join(results, spec["analytic"]), record)
                payload["golden_curve_ph"] = ph
                payload["golden_curve_analytic"] = p
            for k, d in enumerate(spec["mc"]):
                ph, p = formats.read_curves_directory(os.path.join(results, d), record)
                payload["golden_curve_ph"] = ph
                payload["golden_curve_mc_%s" % d] = p

            sub = parse_substate_table(log_text)
            if sub is not None:
                sites, energies, protons, labels = sub
                payload["golden_substate_sites"] = np.array(["%s|%s|%d" % s for s in sites])
                payload["golden_substate_energy"] = np.array(energies)
                payload["golden_substate_protons"] = np.array(protons, dtype=np.int32)
                payload["golden_substate_labels"] = np.array(labels)
            halfpk = parse_halfpk_table(log_text, record)
            if halfpk is not None:
                payload["golden_halfpk"] = halfpk

            gint = find(results, "job.gint")
            inter = find(results, "job.inter")
            if gint and inter:
                grows = formats.read_gintr(gint)
                payload["golden_job_gint"] = np.array([r[4] for r in grows])
                payload["golden_job_protons"] = np.array([r[5] for r in grows], dtype=np.int32)
                n = record.ninstances
                wsym = np.zeros((n, n))
                wraw = np.zeros((n, n))
                index = {(r[0], r[1]): i for i, r in enumerate(grows)}
                for r in formats.read_w(inter):
                    wsym[index[(r[0], r[1])], index[(r[4], r[5])]] = r[8]
                    wraw[index[(r[0], r[1])], index[(r[4], r[5])]] = r[9]
                payload["golden_job_wsym"] = wsym
                payload["golden_job_wraw"] = wraw
                # keep a verbatim head of each table as a reader/writer format fixture
                with open(gint) as handle:
                    payload["golden_job_gint_head"] = np.array("".join(handle.readlines()[:8]))
                with open(inter) as handle:
                    payload["golden_job_inter_head"] = np.array("".join(handle.readlines()[:8]))

            if name == "twosites":
                # verbatim MEAD logs of one instance: parser fixture
                for fn in ("site_HSP.out", "model_HSP.out"):
                    with open(os.path.join(results, "mead", "PRTA", "HIS5", fn)) as handle:
                        payload["golden_mead_%s" % fn.replace(".", "_")] = np.array(handle.read())

            path = os.path.join(OUT, name + ".npz")
            np.savez_compressed(path, **payload)
            print("%-20s nsites=%3d ninst=%3d nstates=%.4g gmicro=%d pairs=%d -> %s (%d KB)" % (
                name, record.nsites, record.ninstances, float(record.nstates), len(gm), len(pairs), os.path.basename(path), os.path.getsize(path) // 1024))

    extract_lysozyme_compare(parameters)

    # lysozyme_traj: per-scan MC energies at pH 7 (stat_mc.dat) and move counters; same model as 'lysozyme'
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(REFERENCE, "tests", "lysozyme_traj", "results.tgz")) as tar:
            tar.extractall(tmp)
        stat = find(tmp, "stat_mc.dat")
        energies = np.loadtxt(stat)
        log = find(tmp, "lyso_mc_traj.out")
        with open(log) as handle:
            text = handle.read()
        counters = []
        for row in text.splitlines():
            m = re.match(r"^\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s*$", row)
            if m:
                counters.append([int(x) for x in m.groups()])
        np.savez_compressed(os.path.join(OUT, "lysozyme_traj.npz"), stat_mc=energies, counters=np.array(counters, dtype=np.int64))
        print("lysozyme_traj        energies=%d counters=%d" % (len(energies), len(counters)))


if __name__ == "__main__":
    main()
