"""One logged Monte Carlo run of lysozyme at pH 7: the reference's ``tests/lysozyme_traj/lyso_mc_traj.py`` from the point
where MEAD has finished.  Prints the move / flip counter table every 500 scans and writes the energy after every production
scan to ``stat_mc.dat``.  Usage: ``python examples/lyso_mc_traj.py [output directory]`` (needs a B200)."""

import os
import sys

import _fixture
from pcetk_b200 import MEADModel, MCModelDefault
from pcetk_b200.log import logFile


def main(directory="lysozyme_traj_out", log=logFile):
    record, _golden = _fixture.load("lysozyme")
    cem = MEADModel(record, log=log)
    cem.Summary(log=log)

    sampling = MCModelDefault()
    cem.DefineMCModel(sampling, log=log)

    cem.CalculateProbabilities(pH=7.0, logFrequency=500, trajectoryFilename=os.path.join(directory, "stat_mc.dat"), log=log)
    cem.SummaryProbabilities(log=log)
    return dict(model=cem, sampler=sampling)


if __name__ == "__main__":
    main(_fixture.output_directory(sys.argv, "lysozyme_traj_out"))
