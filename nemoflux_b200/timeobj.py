"""Time axis lookup -- mirror of nemoflux/timeobj.py (labels only, not on the compute path).

The time variable is the one whose standard_name is 'time' or whose long_name is 'Time axis'
(timeobj.py:8-13).  Mock data from datagen has none; unlike the reference (which then raises in
getTimeAsString, timeobj.py:31-33) the index is used as the label.
"""
from datetime import datetime, timedelta
import re


class TimeObj(object):

    def __init__(self, nc):
        self.timeVarName = ''
        self.timeVar = []
        self.units = ''
        for vName, var in nc.items():
            if getattr(var, 'standard_name', '') == 'time' or getattr(var, 'long_name', '') == 'Time axis':
                self.timeVarName = vName
                self.timeVar = var[:]
                self.units = getattr(var, 'units', '')

    def getValues(self):
        return self.timeVar[:]

    def getSize(self):
        return len(self.timeVar)

    def getTimeAsDate(self, timeIndex):
        """CF 'seconds|hours|days since YYYY-MM-DD ...' -> date; the index itself when there is no time axis"""
        if len(self.timeVar) == 0:
            return timeIndex
        m = re.match(r'\s*(seconds|minutes|hours|days)\s+since\s+(\d+)-(\d+)-(\d+)', str(self.units))
        if not m:
            return timeIndex
        scale = dict(seconds=1., minutes=60., hours=3600., days=86400.)[m.group(1)]
        t0 = datetime(int(m.group(2)), int(m.group(3)), int(m.group(4)))
        return (t0 + timedelta(seconds=float(self.timeVar[timeIndex]) * scale)).date()

    def getTimeAsString(self, timeIndex):
        d = self.getTimeAsDate(timeIndex)
        if isinstance(d, int):
            return f'time index {d}'
        return f'{d.year}-{d.month}-{d.day}'
