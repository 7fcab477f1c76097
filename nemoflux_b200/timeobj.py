"""Time axis labels (not on the compute path).

Interface of nemoflux/timeobj.py: TimeObj(nc) finds the variable whose standard_name is 'time' or whose long_name
is 'Time axis' (timeobj.py:8-13); getValues / getSize / getTimeAsDate / getTimeAsString.  The reference gets
calendar objects from xarray; here the raw numbers are decoded from the CF units ('<unit> since Y-M-D').  Mock
data from datagen has no time axis: the reference raises in getTimeAsString (timeobj.py:31-33), this class labels
the step with its index instead.
"""
import datetime
import re

_SECONDS = {'seconds': 1.0, 'minutes': 60.0, 'hours': 3600.0, 'days': 86400.0}
_SINCE = re.compile(r'\s*(seconds|minutes|hours|days)\s+since\s+(\d+)-(\d+)-(\d+)')


def _is_time_axis(var):
    return getattr(var, 'standard_name', '') == 'time' or getattr(var, 'long_name', '') == 'Time axis'


class TimeObj(object):

    def __init__(self, nc):
        self.timeVarName, self.timeVar, self.units = '', [], ''
        for name, var in nc.items():
            if _is_time_axis(var):          # the last match wins, as in the reference
                self.timeVarName, self.timeVar, self.units = name, var[:], str(getattr(var, 'units', ''))
        since = _SINCE.match(self.units)
        self._origin = datetime.datetime(*(int(since.group(k)) for k in (2, 3, 4))) if since else None
        self._unit_s = _SECONDS[since.group(1)] if since else None

    def getValues(self):
        return self.timeVar[:]

    def getSize(self):
        return len(self.timeVar)

    def getTimeAsDate(self, timeIndex):
        """calendar date of the step, or the index itself when the file has no (decodable) time axis"""
        if self._origin is None or len(self.timeVar) == 0:
            return timeIndex
        offset = datetime.timedelta(seconds=float(self.timeVar[timeIndex]) * self._unit_s)
        return (self._origin + offset).date()

    def getTimeAsString(self, timeIndex):
        when = self.getTimeAsDate(timeIndex)
        if isinstance(when, datetime.date):
            return f'{when.year}-{when.month}-{when.day}'
        return f'time index {when}'
