"""Transect text files (WOCE-style station lists), the input of fluxplot's -i flag.

Interface of the reference reader (nemoflux/latlonreader.py:8-20): ``LatLonReader(path).getLonLats()`` gives an
(n, 2) array of (lon, lat).  A station line reads ``STA  DIST  LAT  LON ...`` with STA a whole number and DIST
a number written with a decimal point; headers, rules and blank lines are anything else.  The file lists the
latitude first -- the points come back longitude first.
"""
import re
import sys

import numpy

_WHOLE = re.compile(r'\d+$')
_DECIMAL = re.compile(r'\d+\.\d+$')
_COORD = re.compile(r'-?\d+\.?\d*')


def stations(lines):
    """(lon, lat) of every station line of an iterable of text lines"""
    for text in lines:
        tok = text.split()
        if len(tok) < 4 or not _WHOLE.match(tok[0]) or not _DECIMAL.match(tok[1]):
            continue
        # the reference's pattern ends after the longitude: LAT must be a clean number, LON may carry a suffix
        lat, lon = _COORD.fullmatch(tok[2]), _COORD.match(tok[3])
        if lat and lon:
            yield float(lon.group()), float(lat.group())


class LatLonReader(object):
    """the reference class: reads the file once, keeps the station list"""

    def __init__(self, filename):
        with open(filename) as handle:
            self.lonLatTargets = list(stations(handle))

    def getLonLats(self):
        return numpy.array(self.lonLatTargets)


if __name__ == '__main__':
    print(LatLonReader(sys.argv[1]).getLonLats())
