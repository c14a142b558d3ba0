"""Protonation states in lysozyme: the reference's ``tests/lysozyme/lysozyme.py`` from the point where MEAD has finished.
Analytic curves over 4 194 304 states per pH value, Monte Carlo curves with the default sampler, pK1/2 table, GMCT
tables and the histidine sed script.  Usage: ``python examples/lysozyme.py [output directory]`` (needs a B200)."""

import os
import sys

import _fixture
from pcetk_b200 import MEADModel, MCModelDefault, TitrationCurves
from pcetk_b200.log import logFile


def main(directory="lysozyme_out", log=logFile):
    record, _golden = _fixture.load("lysozyme")
    cem = MEADModel(record, log=log)
    cem.Summary(log=log)
    cem.SummarySites(log=log)
    cem.WriteGintr(os.path.join(directory, "gintr.dat"))
    cem.WriteW(os.path.join(directory, "W.dat"))

    log.Text("\n*** Calculating titration curves analytically ***\n")
    ca = TitrationCurves(cem, log=log)
    ca.CalculateCurves(log=log)
    ca.WriteCurves(directory=os.path.join(directory, "curves_analytic"), log=log)
    ca.CalculateHalfpKs()
    ca.PrintHalfpKs(log=log)

    log.Text("\n*** Calculating titration curves using in-house MC sampling ***\n")
    sampling = MCModelDefault()
    cem.DefineMCModel(sampling, log=log)
    cmc = TitrationCurves(cem, log=log)
    cmc.CalculateCurves(log=log)
    cmc.WriteCurves(directory=os.path.join(directory, "curves_mc"), log=log)

    cem.CalculateProbabilities(pH=7.0, log=log)
    cem.SummaryProbabilities(log=log)
    cem.SedScript_FromProbabilities(filename=os.path.join(directory, "his_repl.sed"), overwrite=True, log=log)
    return dict(analytic=ca, mc=cmc, model=cem)


if __name__ == "__main__":
    main(_fixture.output_directory(sys.argv, "lysozyme_out"))
