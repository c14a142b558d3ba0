"""TitrationCurves: pH sweeps of protonation probabilities (``ContinuumElectrostatics/TitrationCurves.py``).

Same constructor keywords, ``CalculateCurves`` / ``WriteCurves`` / ``CalculateHalfpKs`` / ``PrintHalfpKs``.  The
reference computes one pH value per call (serially, or in Python threads that share one model,
TitrationCurves.py:93-165); here the whole grid goes to the GPU in one call unless ``forceSerial`` asks
for the reference's per-pH loop.
"""

import os

from .errors import ContinuumElectrostaticsError
from .log import logFile, LogFileActive

_DefaultDirectory = "curves"
_DefaultSampling = .5
_DefaultStart = 0.
_DefaultStop = 14.


class TitrationCurves(object):
    """Titration curves."""

    defaultAttributes = {"curveSampling": _DefaultSampling, "curveStart": _DefaultStart, "curveStop": _DefaultStop, "unfolded": False}

    def __init__(self, meadModel, log=logFile, *arguments, **keywordArguments):
        for (key, value) in self.__class__.defaultAttributes.items():
            setattr(self, key, value)
        for (key, value) in keywordArguments.items():
            setattr(self, key, value)
        if not meadModel.isCalculated:
            raise ContinuumElectrostaticsError("First calculate electrostatic energies.")
        self.owner = meadModel
        self.nsteps = int((self.curveStop - self.curveStart) / self.curveSampling + 1)     # TitrationCurves.py:62
        self.steps = None
        self.halves = None
        self.isHalves = False
        self.isCalculated = False

    def pHValues(self):
        return [self.curveStart + step * self.curveSampling for step in range(self.nsteps)]

    def CalculateCurves(self, forceSerial=False, printTable=False, log=logFile):
        """Calculate titration curves.  ``forceSerial=True`` reproduces the reference's one-call-per-pH loop
        (TitrationCurves.py:119-121); the default sends the whole grid in one batch."""
        if not self.isCalculated:
            owner = self.owner
            restore = False
            if owner.isProbability:
                probabilities = owner.energyModel.GetProbabilities()      # TitrationCurves.py:70-89, in bulk
                restore = True
            values = self.pHValues()
            if LogFileActive(log):
                log.Text("\nStarting %s run.\n" % ("serial" if forceSerial else "batched"))
            if forceSerial:
                steps = [owner.CalculateProbabilities(pH=pH, log=None, isCalculateCurves=True, unfolded=self.unfolded) for pH in values]
            else:
                table = owner.CalculateProbabilitiesBatch(values, unfolded=self.unfolded)
                steps = [owner._Gather(row) for row in table]
            if LogFileActive(log):
                if printTable:
                    log.Table([("%10d" % step, "%10.2f" % pH) for step, pH in enumerate(values)], title="Step / pH")
                log.Text("\nCalculating titration curves complete.\n")
            if restore:
                owner.energyModel.SetProbabilities(probabilities)
            else:
                owner.isProbability = False
            self.isCalculated = True
            self.steps = steps

    def WriteCurves(self, directory=_DefaultDirectory, log=logFile):
        """One ``<site.label>_<instance.label>.dat`` per instance, lines ``%f %f`` (TitrationCurves.py:169-194)."""
        if self.isCalculated:
            if not os.path.exists(directory):
                try:
                    os.makedirs(directory)
                except OSError:
                    raise ContinuumElectrostaticsError("Cannot create directory %s" % directory)
            for site in self.owner.sites:
                for instance in site.instances:
                    lines = ["%f %f\n" % (self.curveStart + step * self.curveSampling, self.steps[step][site.siteIndex][instance.instIndex])
                             for step in range(self.nsteps)]
                    with open(os.path.join(directory, "%s_%s.dat" % (site.label, instance.label)), "w") as handle:
                        handle.writelines(lines)
            if LogFileActive(log):
                log.Text("\nWriting curve files complete.\n")

    def CalculateHalfpKs(self):
        """pK1/2 by linear interpolation wherever a curve crosses 0.5 (TitrationCurves.py:198-222)."""
        if self.isCalculated:
            halves = []
            for site in self.owner.sites:
                instances = []
                for instance in site.instances:
                    pa = self.steps[0][site.siteIndex][instance.instIndex]
                    pKs = []
                    for step in range(1, self.nsteps):
                        pb = self.steps[step][site.siteIndex][instance.instIndex]
                        if ((pa - .5) * (pb - .5)) < 0.:
                            a = self.curveStart + (step - 1) * self.curveSampling
                            b = self.curveStart + step * self.curveSampling
                            pKs.append(a + (.5 - pa) * (b - a) / (pb - pa))
                        pa = pb
                    instances.append(pKs)
                halves.append(instances)
            self.halves = halves
            self.isHalves = True

    def _GetEntry(self, site, decimalPlaces=2):
        entry = ""
        if self.isHalves:
            for instance in site.instances:
                pKs = self.halves[site.siteIndex][instance.instIndex]
                entry = "%s%7s" % (entry, instance.label)
                if len(pKs) < 1:
                    entry = "%s%7s" % (entry, "n/a")
                else:
                    for pK in pKs:
                        entry = ("%%s%%7.%df" % decimalPlaces) % (entry, pK)
        return entry

    def PrintHalfpKs(self, decimalPlaces=2, sortSites=False, log=logFile):
        if LogFileActive(log) and self.isCalculated:
            if not self.isHalves:
                self.CalculateHalfpKs()
            table = [[site.segName, site.resName, site.resSerial, self._GetEntry(site, decimalPlaces)] for site in self.owner.sites]
            if sortSites:
                table.sort(key=lambda k: (k[0], k[1], k[2]))
            log.Table([("%6s" % a, "%6s" % b, "%6d" % c, d) for a, b, c, d in table], title="pK1/2 values of instances")
