"""Substate: energies of all protonation forms of a few selected sites around the most probable state
(``ContinuumElectrostatics/Substate.py``).  The reference makes one C call per substate from a Python loop
(:91-119); here every substate goes into one kernel-1 batch.
"""

import numpy as np

from .errors import ContinuumElectrostaticsError
from .log import logFile, LogFileActive
from .StateVector import StateVector


def StateVector_FromProbabilities(meadModel):
    """State vector of the most probable instance per site (Substate.py:232-246; ties -> highest index, as ``max`` of pairs)."""
    if not meadModel.isProbability:
        raise ContinuumElectrostaticsError("First calculate probabilities.")
    vector = StateVector(meadModel)
    for siteIndex, site in enumerate(meadModel.sites):
        pairs = [(instance.probability, instanceIndex) for instanceIndex, instance in enumerate(site.instances)]
        _probability, instanceIndex = max(pairs)
        vector[siteIndex] = instanceIndex
    return vector


class Substate(object):
    """Substate of a protonation state."""

    def __init__(self, meadModel, selectedSites, pH=7.0, log=logFile):
        """|selectedSites| is a sequence of (segmentName, residueName, residueSerial)."""
        pairs, indicesOfSites = [], []
        for selectedSegment, selectedResidueName, selectedResidueSerial in selectedSites:
            for siteIndex, site in enumerate(meadModel.sites):
                if site.segName == selectedSegment and site.resSerial == selectedResidueSerial:
                    indicesOfSites.append(siteIndex)
                    break
            else:
                raise ContinuumElectrostaticsError("Site %s %s %d not found." % (selectedSegment, selectedResidueName, selectedResidueSerial))
            pairs.append([selectedSegment, selectedResidueSerial])
        self.indicesOfSites = indicesOfSites
        self.isCalculated = False
        self.substates = None
        self.owner = meadModel
        self.pH = pH
        self.vector = self._DetermineLowestEnergyVector()
        self.vector.DefineSubstate(pairs)
        if LogFileActive(log):
            nsites = len(indicesOfSites)
            log.Text("\nSubstate is initialized with %d site%s.\n" % (nsites, "s" if nsites > 1 else ""))

    def _DetermineLowestEnergyVector(self):
        owner = self.owner
        backup = owner.energyModel.GetProbabilities() if owner.isProbability else None
        owner.CalculateProbabilities(pH=self.pH, log=None)
        vector = StateVector_FromProbabilities(owner)
        if backup is not None:
            owner.energyModel.SetProbabilities(backup)
        else:
            owner.isProbability = False       # (the reference misspells this attribute, Substate.py:63)
        return vector

    def CalculateSubstateEnergies(self, log=logFile):
        """Enumerate the substate (the IncrementSubstate loop of Substate.py:99-108, run natively) and evaluate all of it in
        one kernel-1 batch (Substate.py:91-119)."""
        if not self.isCalculated:
            # every substate as a row of global instance indices, in IncrementSubstate order (native odometer)
            states = self.vector.SubstateRows()
            firsts = np.array([self.owner.sites[siteIndex].instances[0]._instIndexGlobal for siteIndex in self.indicesOfSites], dtype=np.int32)
            locals_ = (states[:, self.indicesOfSites] - firsts).tolist()
            energies = self.owner.energyModel.CalculateMicrostateEnergies(states, pH=float(self.pH))
            substates = [[float(energy), indices] for energy, indices in zip(energies, locals_)]
            substates.sort()
            self.substates = substates
            self.zeroEnergy = substates[0][0]
            self.isCalculated = True
            if LogFileActive(log):
                log.Text("\nCalculating substate energies at pH=%.1f complete.\n" % self.pH)

    def Summary(self, relativeEnergy=True, roundCharge=True, title="", log=logFile):
        """Table of substate energies (Substate.py:122-168; the Charge column needs atomic charges, which the
        precomputed-energy input does not carry, so it is omitted)."""
        if self.isCalculated and LogFileActive(log):
            owner = self.owner
            header = ["State", "Gmicro", "Protons"] + ["%s %s %d" % (owner.sites[i].segName, owner.sites[i].resName, owner.sites[i].resSerial)
                                                       for i in self.indicesOfSites]
            rows = [header]
            for stateIndex, (energy, indicesOfInstances) in enumerate(self.substates):
                if relativeEnergy:
                    energy = energy - self.zeroEnergy
                nprotons, labels = 0, []
                for siteIndex, instanceIndex in zip(self.indicesOfSites, indicesOfInstances):
                    instance = owner.sites[siteIndex].instances[instanceIndex]
                    nprotons += instance.protons
                    labels.append(instance.label)
                rows.append(["%6d" % (stateIndex + 1), "%9.2f" % energy, "%d" % nprotons] + labels)
            log.Table(rows, title=title or None)


    def Summary_ToLatex(self, filename="table.tex", relativeEnergy=True, includeSegment=False, translateToLatex=None):
        """Write the substate table as a LaTeX ``tabular`` (Substate.py:171-220; same columns and cell formats)."""
        symbols = {"ASP": {"p": "0", "d": "($-$)"},
                   "GLU": {"p": "0", "d": "($-$)"},
                   "HIS": {"HSP": "$\\epsilon$, $\\delta$(+)", "HSE": "$\\epsilon$", "HSD": "$\\delta$", "fd": "($-$)"}}
        if translateToLatex:
            symbols.update(translateToLatex)
        if not self.isCalculated:
            return
        sites = [self.owner.sites[siteIndex] for siteIndex in self.indicesOfSites]
        titles = ["%s%s%d" % ((site.segName + " ") if includeSegment else "", site.resName.capitalize(), site.resSerial) for site in sites]
        lines = ["\\begin{tabular}{@{\\extracolsep{2mm}}cc%sc}" % ("l" * len(sites)),
                 "State & $\\Delta E$ (kcal/mol) & " + "".join(" %s &" % title for title in titles) + " No. of protons \\\\",
                 "\\hline\\noalign{\\smallskip}"]
        for count, (energy, indicesOfInstances) in enumerate(self.substates, 1):
            if relativeEnergy:
                energy = energy - self.zeroEnergy
            cells, nprotons = [], 0
            for site, instanceIndex in zip(sites, indicesOfInstances):
                instance = site.instances[instanceIndex]
                cells.append("  %-26s &" % symbols.get(site.resName, {}).get(instance.label, instance.label))
                nprotons += instance.protons
            lines.append("%3d & %5.1f & " % (count, energy) + "".join(cells) + "  %1d \\\\" % nprotons)
        lines.extend(["\\hline\\noalign{\\smallskip}", "\\end{tabular}"])
        with open(filename, "w") as handle:
            handle.write("\n".join(lines) + "\n")


class MEADSubstate(Substate):
    """A class to maintain backward compatibility (Substate.py:224-226)."""
    pass
