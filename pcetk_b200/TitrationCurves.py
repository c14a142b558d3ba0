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
        settings = dict(self.defaultAttributes)
        settings.update(keywordArguments)
        vars(self).update(settings)
        if not meadModel.isCalculated:
            raise ContinuumElectrostaticsError("First calculate electrostatic energies.")
        self.owner = meadModel
        self.nsteps = int((self.curveStop - self.curveStart) / self.curveSampling + 1)     # grid convention of TitrationCurves.py:62
        self.steps = self.halves = None
        self.isHalves = self.isCalculated = False

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
                steps = owner._GatherRows(table)
                self._table = table
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

    def _CurveTable(self):
        """The calculated curves as one array [nsteps, ninstances] (global instance order) and the pH grid."""
        import numpy as np

        table = getattr(self, "_table", None)          # kept by the batched path; rebuilt from the nested lists otherwise
        if table is None or len(table) != len(self.steps):
            table = np.array([[value for site in step for value in site] for step in self.steps], dtype=np.float64)
        return table, self.curveStart + self.curveSampling * np.arange(self.nsteps)

    def CalculateHalfpKs(self):
        """pK1/2 of every instance: the pH values where its curve crosses 0.5, by linear interpolation between the two
        grid points around the crossing (same definition as TitrationCurves.py:198-222; vectorised over the grid)."""
        if not self.isCalculated:
            return
        import numpy as np

        table, grid = self._CurveTable()
        centred = table - 0.5
        crossing = centred[:-1] * centred[1:] < 0.0                      # [nsteps - 1, ninstances]
        columns, left = np.nonzero(crossing.T)                           # every crossing, ordered by instance, then by pH
        pa, pb = table[left, columns], table[left + 1, columns]
        values = (grid[left] + (0.5 - pa) * (grid[left + 1] - grid[left]) / (pb - pa)).tolist()
        bounds = np.searchsorted(columns, np.arange(table.shape[1] + 1)).tolist()
        self.halves, column = [], 0
        for site in self.owner.sites:
            count = len(site.instances)
            self.halves.append([values[bounds[k]:bounds[k + 1]] for k in range(column, column + count)])
            column += count
        self.isHalves = True

    def _GetEntry(self, site, decimalPlaces=2):
        """One table cell: label and pK1/2 value(s) of every instance of the site, 7 characters per field."""
        if not self.isHalves:
            return ""
        number = "%%7.%df" % decimalPlaces
        fields = []
        for instance in site.instances:
            values = self.halves[site.siteIndex][instance.instIndex]
            fields.append("%7s" % instance.label)
            fields.extend([number % value for value in values] or ["%7s" % "n/a"])
        return "".join(fields)

    def PrintHalfpKs(self, decimalPlaces=2, sortSites=False, log=logFile):
        if LogFileActive(log) and self.isCalculated:
            if not self.isHalves:
                self.CalculateHalfpKs()
            table = [[site.segName, site.resName, site.resSerial, self._GetEntry(site, decimalPlaces)] for site in self.owner.sites]
            if sortSites:
                table.sort(key=lambda k: (k[0], k[1], k[2]))
            log.Table([("%6s" % a, "%6s" % b, "%6d" % c, d) for a, b, c, d in table], title="pK1/2 values of instances")
