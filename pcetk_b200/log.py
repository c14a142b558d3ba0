"""A minimal stand-in for pDynamo's ``logFile`` (pCore, not vendored by the reference).

The reference prints through ``log.Text``, ``log.GetTable`` and ``log.GetSummary``; only plain text
output is reproduced here.  ``LogFileActive(log)`` is False for ``None``, as in pCore.
"""

import sys


class TextLog(object):
    def __init__(self, stream=None):
        self.stream = stream if stream is not None else sys.stdout

    def Text(self, text):
        self.stream.write(text)

    def Table(self, rows, title=None):
        widths = [max(len(str(row[k])) for row in rows) for k in range(len(rows[0]))] if rows else []
        rule = "-" * (sum(widths) + 2 * len(widths))
        if title:
            self.stream.write(title.center(len(rule)) + "\n")
        self.stream.write(rule + "\n")
        for row in rows:
            self.stream.write("  ".join(str(c).rjust(w) for c, w in zip(row, widths)) + "\n")
        self.stream.write(rule + "\n")


logFile = TextLog()


def LogFileActive(log):
    return log is not None
