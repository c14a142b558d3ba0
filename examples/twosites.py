"""Two titratable sites in a hypothetical peptide: the reference's ``tests/twosites/twosites.py`` from the point where MEAD
has finished.  All eight microstate energies at pH 7, probabilities at pH 7 by enumeration and by Monte Carlo, and both
sets of titration curves.  Usage: ``python examples/twosites.py [output directory]`` (needs a B200)."""

import os
import sys

import _fixture
from pcetk_b200 import MEADModel, StateVector, MCModelDefault, TitrationCurves
from pcetk_b200.log import logFile


def main(directory="twosites_out", log=logFile):
    record, _golden = _fixture.load("twosites")
    cem = MEADModel(record, log=log)
    cem.Summary(log=log)
    cem.SummarySites(log=log)

    log.Text("\n*** Calculating microstate energies of all states at pH=7 ***\n")
    statevector = StateVector(cem)
    energies = []
    increment = True
    while increment:
        Gmicro = cem.CalculateMicrostateEnergy(statevector, pH=7.0)
        statevector.Print(title="Gmicro = %f" % Gmicro, log=log)
        energies.append(Gmicro)
        increment = statevector.Increment()

    log.Text("\n*** Calculating protonation probabilities at pH=7 analytically ***\n")
    cem.CalculateProbabilities(pH=7.0, log=log)
    cem.SummaryProbabilities(log=log)

    log.Text("\n*** Calculating protonation probabilities at pH=7 using in-house MC sampling ***\n")
    mc = MCModelDefault(nprod=30000)
    cem.DefineMCModel(mc, log=log)
    cem.CalculateProbabilities(log=log)
    cem.SummaryProbabilities(log=log)

    log.Text("\n*** Calculating titration curves using in-house MC sampling ***\n")
    cmc = TitrationCurves(cem, curveSampling=0.5, log=log)
    cmc.CalculateCurves(log=log)
    cmc.WriteCurves(directory=os.path.join(directory, "curves_mc"), log=log)

    log.Text("\n*** Calculating titration curves analytically ***\n")
    cem.DefineMCModel(None)
    ca = TitrationCurves(cem, curveSampling=0.5, log=log)
    ca.CalculateCurves(log=log)
    ca.WriteCurves(directory=os.path.join(directory, "curves_analytic"), log=log)
    return dict(energies=energies, mc=cmc, analytic=ca, model=cem)


if __name__ == "__main__":
    main(_fixture.output_directory(sys.argv, "twosites_out"))
