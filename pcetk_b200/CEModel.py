"""CEModel: the continuum-electrostatics model as the host of the microstate path.

Keeps the call surface of ``ContinuumElectrostatics/CEModel.py`` that the hot path touches --
``nsites``/``ninstances`` (:52-66), ``DefineMCModel`` (:442-456), ``CalculateMicrostateEnergy`` (:460-462),
``CalculateProbabilities`` (:466-499), ``SummarySites`` / ``SummaryProbabilities`` (:192-274),
``SedScript_FromProbabilities`` (:278-331), ``PrintInteractions`` (:335-359), ``WriteW``/``WriteGintr`` (:363-438) -- with ``Site`` and ``Instance`` objects carrying the same attributes (``Site.py``,
``Instance.py:37-75``).

The reference builds a model from a pDynamo ``System`` and runs MEAD; both are out of scope (SURVEY.md
section 2, rows 8-10).  Here a model is built from the *inputs of the hot path* -- a
:class:`~pcetk_b200.formats.ModelRecord` (Gintr, protons, raw W per instance) obtained from cached MEAD
logs, GMCT tables, or a synthetic generator.
"""

import os

import numpy as np

from .constants import DEFAULT_TEMPERATURE
from .errors import ContinuumElectrostaticsError
from .EnergyModel import EnergyModel
from .MCModelDefault import MCModelDefault
from .formats import ModelRecord, format_gintr_lines, format_w_lines, model_from_gmct_tables, model_from_mead_directory
from .log import logFile, LogFileActive


class Instance(object):
    """An instance (protonation form) of a titratable site; energies live in the energy model (Instance.py:37-75)."""

    def __init__(self, **keywordArguments):
        for (key, val) in keywordArguments.items():
            setattr(self, key, val)

    def _GetEnergyModel(self):
        if hasattr(self, "parent"):
            return self.parent.parent.energyModel
        raise ContinuumElectrostaticsError("Energy model is undefined.")

    @property
    def Gmodel(self):
        return self._GetEnergyModel().GetGmodel(self._instIndexGlobal)

    @Gmodel.setter
    def Gmodel(self, value):
        self._GetEnergyModel().SetGmodel(self._instIndexGlobal, value)

    @property
    def Gintr(self):
        return self._GetEnergyModel().GetGintr(self._instIndexGlobal)

    @Gintr.setter
    def Gintr(self, value):
        self._GetEnergyModel().SetGintr(self._instIndexGlobal, value)

    @property
    def protons(self):
        return self._GetEnergyModel().GetProtons(self._instIndexGlobal)

    @protons.setter
    def protons(self, value):
        self._GetEnergyModel().SetProtons(self._instIndexGlobal, value)

    @property
    def probability(self):
        return self._GetEnergyModel().GetProbability(self._instIndexGlobal)

    @probability.setter
    def probability(self, value):
        self._GetEnergyModel().SetProbability(self._instIndexGlobal, value)

    @property
    def interactions(self):
        energyModel = self._GetEnergyModel()
        return [energyModel.GetInteractionSymmetric(self._instIndexGlobal, other) for other in range(self.parent.parent.ninstances)]


class Site(object):
    """A titratable site (Site.py): ``segName``, ``resName``, ``resSerial``, ``siteIndex``, ``instances``."""

    def __init__(self, parent, siteIndex, segName, resName, resSerial):
        self.parent = parent
        self.siteIndex = siteIndex
        self.segName = segName
        self.resName = resName
        self.resSerial = resSerial
        self.instances = []

    @property
    def ninstances(self):
        return len(self.instances)

    @property
    def label(self):
        return "%s_%s%d" % (self.segName, self.resName, self.resSerial)

    def GetMostProbableInstance(self):
        best = max(self.instances, key=lambda instance: instance.probability)
        return (best.probability, best.instIndex, best.label)


class CEModel(object):
    """Continuum electrostatic model reduced to what the microstate path needs."""

    defaultAttributes = {"temperature": DEFAULT_TEMPERATURE, "nthreads": 1, "device": 0}

    def __init__(self, record=None, log=logFile, **keywordArguments):
        attributes = self.__class__.defaultAttributes
        for (key, val) in attributes.items():
            setattr(self, key, val)
        for (key, val) in keywordArguments.items():
            if key not in attributes:
                raise ContinuumElectrostaticsError("Unknown attribute: %s" % key)
            setattr(self, key, val)
        for attribute in ("isInitialized", "isFilesWritten", "isCalculated", "isProbability"):
            setattr(self, attribute, False)
        if record is not None:
            self.InitializeFromRecord(record, log=log)

    @property
    def label(self):
        return "B200"

    @property
    def nsites(self):
        return len(self.sites) if hasattr(self, "sites") else 0

    @property
    def ninstances(self):
        return sum(site.ninstances for site in self.sites) if hasattr(self, "sites") else 0

    # ---- construction -------------------------------------------------------------------------------------
    def InitializeFromRecord(self, record, asymmetricTolerance=0.05, log=logFile):
        """Replaces ``Initialize`` + ``CalculateElectrostaticEnergies`` (CEModel.py:114-139, CEModelMEAD.py:91-184)
        for precomputed energies: create sites/instances, load Gmodel/protons/Gintr/interactions, check symmetry,
        symmetrise."""
        record.validate()
        self.temperature = record.temperature
        self.record = record
        self.sites = []
        labels = record.site_labels or [("PRTA", "SIT", k + 1) for k in range(record.nsites)]
        inst_labels = record.inst_labels or ["i%d" % k for k in range(record.ninstances)]
        self.energyModel = EnergyModel(self, record.nsites, record.ninstances, device=self.device)
        for siteIndex, (segName, resName, resSerial) in enumerate(labels):
            site = Site(self, siteIndex, segName, resName, resSerial)
            for instIndex, g in enumerate(range(int(record.site_first[siteIndex]), int(record.site_last[siteIndex]) + 1)):
                site.instances.append(Instance(parent=site, instIndex=instIndex, _instIndexGlobal=g, label=inst_labels[g]))
            self.sites.append(site)
        self.energyModel.Load(gintr=record.gintr, gmodel=record.gmodel, protons=record.protons, interactions=record.interactions)
        self.energyModel.Initialize()
        self.isInitialized = True
        isSymmetric, maxDeviation = self.energyModel.CheckIfSymmetric(tolerance=asymmetricTolerance)
        if LogFileActive(log):
            if isSymmetric:
                log.Text("\nInteractions are symmetric within the given tolerance (%0.4f kcal/mol).\n" % asymmetricTolerance)
            else:
                log.Text("\nWARNING: Maximum deviation of interactions is %0.4f kcal/mol.\n" % maxDeviation)
        self.energyModel.SymmetrizeInteractions(log=log)
        self.isCalculated = True

    @classmethod
    def FromMEADOutputs(cls, name, logText, meadDirectory, parameters, temperature=DEFAULT_TEMPERATURE, log=logFile, **keywordArguments):
        """Model from cached MEAD solver logs (what a rerun of the reference reads, InstanceMEAD.py:36,73)."""
        return cls(model_from_mead_directory(name, logText, meadDirectory, parameters, temperature), log=log, **keywordArguments)

    @classmethod
    def FromGMCTTables(cls, name, gintPath, interPath, temperature=DEFAULT_TEMPERATURE, log=logFile, **keywordArguments):
        """Model from ``WriteGintr`` / ``WriteW`` tables (CEModel.py:363-438)."""
        return cls(model_from_gmct_tables(name, gintPath, interPath, temperature), log=log, **keywordArguments)

    # ---- hot-path surface ---------------------------------------------------------------------------------
    def DefineMCModel(self, sampler, log=logFile):
        """Assign a Monte Carlo model (CEModel.py:442-456).  ``None`` detaches the sampler, as the manual says
        (doc/Pcetk_manual.tex:602-607); in the reference's code ``None`` is a silent no-op."""
        if sampler is None:
            if hasattr(self, "sampler"):
                del self.sampler
            return
        if not isinstance(sampler, MCModelDefault):
            raise ContinuumElectrostaticsError("Cannot define MC model.")
        self.sampler = sampler
        self.sampler.Initialize(self)
        self.sampler.PrintPairs(log=log)
        self.sampler.Summary(log=log)

    def CalculateMicrostateEnergy(self, stateVector, pH=7.0):
        return self.energyModel.CalculateMicrostateEnergy(stateVector, pH=pH)

    def _Gather(self, row=None):
        """Probabilities of one pH value as the reference's nested lists [site][instance] (CEModel.py:491-497)."""
        if row is None:
            row = self.energyModel.GetProbabilities()
        return self._GatherRows([row])[0]

    def _GatherRows(self, table):
        """The same for many pH values at once: rows of p[npH, ninst] -> [step][site][instance] (one bulk conversion, then slices)."""
        spans = getattr(self, "_siteSpans", None)
        if spans is None or len(spans) != len(self.sites):
            spans = []
            for site in self.sites:
                indices = [instance._instIndexGlobal for instance in site.instances]
                contiguous = indices == list(range(indices[0], indices[0] + len(indices)))
                spans.append((indices[0], indices[0] + len(indices)) if contiguous else indices)
            self._siteSpans = spans
        rows = [list(map(float, row)) for row in table] if not hasattr(table, "tolist") else table.tolist()
        if all(isinstance(span, tuple) for span in spans):
            return [[row[first:last] for first, last in spans] for row in rows]
        return [[row[span[0]:span[1]] if isinstance(span, tuple) else [row[k] for k in span] for span in spans] for row in rows]

    def CalculateProbabilities(self, pH=7.0, unfolded=False, isCalculateCurves=False, logFrequency=-1, trajectoryFilename="", log=logFile):
        """Calculate probabilities (CEModel.py:466-499): the sampler if one is attached, else exhaustive enumeration."""
        nstates = -1
        sites = None
        hasSampler = hasattr(self, "sampler")
        if hasSampler and unfolded:
            raise ContinuumElectrostaticsError("Monte Carlo sampling of unfolded proteins unsupported.")
        elif hasSampler:
            self.sampler.CalculateOwnerProbabilities(pH=pH, logFrequency=logFrequency, trajectoryFilename=trajectoryFilename, log=log)
        elif unfolded:
            if trajectoryFilename != "":
                raise ContinuumElectrostaticsError("Writing trajectories of unfolded proteins unsupported.")
            nstates = self.energyModel.CalculateProbabilitiesAnalyticallyUnfolded(pH=pH)
        else:
            if trajectoryFilename != "":
                raise ContinuumElectrostaticsError("Writing trajectories unsupported.")
            nstates = self.energyModel.CalculateProbabilitiesAnalytically(pH=pH)
        if isCalculateCurves:
            sites = self._Gather()
        if nstates > 0 and LogFileActive(log):
            log.Text("\nCalculated %d protonation states.\n" % nstates)
        self.isProbability = True
        return sites

    def CalculateProbabilitiesBatch(self, pHs, unfolded=False):
        """All pH values in one GPU call -> array p[npH, ninst].  The per-pH loop of
        ``TitrationCurves.CalculateCurves`` (TitrationCurves.py:119-121) becomes the pH axis of the launch."""
        if hasattr(self, "sampler"):
            if unfolded:
                raise ContinuumElectrostaticsError("Monte Carlo sampling of unfolded proteins unsupported.")
            p = self.sampler.CalculateOwnerProbabilitiesBatch(pHs)
        else:
            p = self.energyModel.CalculateProbabilitiesAnalyticallyBatch(pHs, unfolded=unfolded)
        self.isProbability = True
        return p

    # ---- reports / files ------------------------------------------------------------------------------------
    def Summary(self, log=logFile):
        if LogFileActive(log):
            log.Table([("Temperature", "%g" % self.temperature), ("Initialized", str(self.isInitialized)),
                       ("Calculated", str(self.isCalculated)), ("Calculated Prob.", str(self.isProbability)),
                       ("Number Of Sites", "%s" % self.nsites), ("Number Of Instances", "%s" % self.ninstances)],
                      title="Continuum Electrostatic Model (%s)" % self.label)

    def SummaryProbabilities(self, reportOnlyUnusual=False, maxProbThreshold=0.75, log=logFile):
        """List probabilities of occurrence of instances (CEModel.py:220-274): '*' marks the most probable one."""
        unusualProtonations = {"HIS": ("HSE", "HSD", "fd"), "ARG": ("d", ), "ASP": ("p", ), "CYS": ("d", ),
                               "GLU": ("p", ), "LYS": ("d", ), "TYR": ("d", )}
        if LogFileActive(log) and self.isProbability:
            rows = []
            for site in self.sites:
                maxProb, maxIndex, maxLabel = site.GetMostProbableInstance()
                skipSite = False
                if reportOnlyUnusual and site.resName in unusualProtonations and maxLabel not in unusualProtonations[site.resName]:
                    skipSite = True
                if maxProb < maxProbThreshold:
                    skipSite = False
                if skipSite:
                    continue
                row = ["%6s" % site.segName, "%6s" % site.resName, "%6d" % site.resSerial]
                for instance in site.instances:
                    row.append("%8s" % (("*%s" % instance.label) if instance.instIndex == maxIndex else instance.label))
                    row.append("%8.4f" % instance.probability)
                rows.append(row)
            width = max(len(r) for r in rows) if rows else 0
            log.Table([r + [""] * (width - len(r)) for r in rows], title="Probabilities of instances")

    def SummarySites(self, log=logFile):
        """List titratable sites (CEModel.py:192-216).  The Center columns need atomic coordinates, which the
        precomputed-energy input does not carry; they are printed only when the record supplies ``site_centers``."""
        if LogFileActive(log) and self.isInitialized:
            centers = getattr(self.record, "site_centers", None)
            rows = [["SiteID", "Segment", "Residue", "Serial", "Instances"] + (["x", "y", "z"] if centers is not None else [])]
            for site in self.sites:
                row = ["%d" % (site.siteIndex + 1), "%s" % site.segName, "%s" % site.resName, "%d" % site.resSerial, "%d" % site.ninstances]
                if centers is not None:
                    row.extend("%10.3f" % c for c in centers[site.siteIndex])
                rows.append(row)
            log.Table(rows, title="Titratable sites")

    def PrintInteractions(self, log=logFile):
        """Print the symmetrised interaction matrix in a readable form (CEModel.py:335-359); readable only for a
        few sites with two instances each, as the reference says."""
        if self.isCalculated and LogFileActive(log):
            wsym = self.energyModel.GetSymmetricMatrix()
            log.Text("%16s" % "" + "".join("%16s" % site.label.center(8) for site in self.sites) + "\n")
            log.Text("%16s" % "" + "".join("%8s" % instance.label.center(8) for site in self.sites for instance in site.instances) + "\n")
            for site in self.sites:
                for instance in site.instances:
                    row = wsym[instance._instIndexGlobal]
                    log.Text("%16s" % (site.label + " " + instance.label) + "".join("%8.2f" % w for w in row) + "\n")

    def SedScript_FromProbabilities(self, filename="his_repl.sed", overwrite=False, putPath=False, log=logFile):
        """Write a sed script that renames histidines in a PDB file to their most probable tautomer, with CHARMM
        patch hints for unusual protonations (CEModel.py:278-331; same line formats)."""
        if not self.isProbability:
            raise ContinuumElectrostaticsError("First calculate probabilities.")
        if not overwrite and os.path.exists(filename):
            if LogFileActive(log):
                log.Text("\nFile %s already exists, skipping.\n" % filename)
            return
        unusualProtonations = {"ARG": ("d", ), "ASP": ("p", ), "CYS": ("d", ), "GLU": ("p", ), "LYS": ("d", ), "TYR": ("d", )}
        translateLabels = {"p": "protonated", "d": "deprotonated"}
        lines = ["# %s\n" % os.path.abspath(filename)] if putPath else []
        warnings = []
        histidines = ("HIS", "HSP")
        for site in self.sites:
            if site.resName in histidines:
                value, _index, label = site.GetMostProbableInstance()
                lines.append("/H.. .%4d/  s/H../%3s/  # %.4f\n" % (site.resSerial, label, value))
        for site in self.sites:
            if site.resName in histidines:
                continue
            value, _index, label = site.GetMostProbableInstance()
            if site.resName not in unusualProtonations:
                warnings.append("# Unknown residue: %4s %3s %4d (most probable instance is \"%s\" with probability of %.4f)\n" % (
                    site.segName, site.resName, site.resSerial, label, value))
            elif label in unusualProtonations[site.resName]:
                if site.resName in ("ASP", "GLU"):
                    lines.append("# patch %sP %4s %4d setup  ! %.4f\n" % (site.resName, site.segName, site.resSerial, value))
                else:
                    warnings.append("# Warning: %4s %3s %4d is %s with probability of %.4f\n" % (
                        site.segName, site.resName, site.resSerial, translateLabels[label], value))
        with open(filename, "w") as handle:
            handle.writelines(lines + warnings)
        if LogFileActive(log):
            log.Text("\nWrote file: %s\n" % filename)

    def _InstanceRows(self):
        return [(site.siteIndex, instance.instIndex, instance.label) for site in self.sites for instance in site.instances]

    def WriteGintr(self, filename="gintr.dat", precision=4, log=logFile):
        """Write a GMCT-compatible file of intrinsic energies (CEModel.py:405-438)."""
        if self.isCalculated:
            rows = [(s, i, lab, self.energyModel.GetGintr(g), self.energyModel.GetProtons(g))
                    for g, (s, i, lab) in enumerate(self._InstanceRows())]
            with open(filename, "w") as handle:
                handle.writelines(format_gintr_lines([site.label for site in self.sites], rows, precision))

    def WriteW(self, filename="W.dat", precision=4, log=logFile):
        """Write a GMCT-compatible matrix of interactions (CEModel.py:363-402)."""
        if self.isCalculated:
            n = self.ninstances
            wsym = self.energyModel.GetSymmetricMatrix()
            wraw = np.array([[self.energyModel.GetInteraction(a, b) for b in range(n)] for a in range(n)])
            with open(filename, "w") as handle:
                handle.writelines(format_w_lines([site.label for site in self.sites], self._InstanceRows(), wsym, wraw, precision))


class MEADModel(CEModel):
    """Name kept for scripts written against the reference (``from ContinuumElectrostatics import MEADModel``)."""

    @property
    def label(self):
        return "MEAD (precomputed)"
