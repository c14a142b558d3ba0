"""Protonation states in IFP2.0 (73 sites, histidines with four forms, a five-form biliverdin ligand, 49 double-move pairs):
the reference's ``tests/ifp20/ifp20.py`` from the point where MEAD has finished.  Monte Carlo titration curves and the
80-row substate table of HIS260 and the three ligand sites.  Usage: ``python examples/ifp20.py [output directory]``
(needs a B200)."""

import os
import sys

import _fixture
from pcetk_b200 import MEADModel, MEADSubstate, MCModelDefault, TitrationCurves
from pcetk_b200.log import logFile


def main(directory="ifp20_out", log=logFile):
    record, _golden = _fixture.load("ifp20")
    mead = MEADModel(record, log=log)
    mead.Summary(log=log)

    sampling = MCModelDefault()
    mead.DefineMCModel(sampling, log=log)

    tc = TitrationCurves(mead, log=log)
    tc.CalculateCurves(log=log)
    tc.WriteCurves(directory=os.path.join(directory, "curves"), log=log)

    selection = (("PRTA", "HIS", 260), ("CHRO", "BLF", 1), ("CHRO", "ACB", 2), ("CHRO", "ACC", 3))
    substate = MEADSubstate(mead, selection, pH=7.0, log=log)
    substate.CalculateSubstateEnergies(log=log)
    substate.Summary(log=log)
    substate.Summary_ToLatex(filename=os.path.join(directory, "table.tex"))
    return dict(curves=tc, substate=substate, model=mead)


if __name__ == "__main__":
    main(_fixture.output_directory(sys.argv, "ifp20_out"))
