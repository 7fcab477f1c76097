"""Transect text files (WOCE station lists) -- mirror of nemoflux/latlonreader.py.

A data line is ``STA(int)  DIST(float with a '.')  LAT  LON ...``; everything else (headers, rules) is
skipped.  NB the file stores latitude first, the points come back as (lon, lat) (latlonreader.py:16-17).
"""
import re
import sys

import numpy

PAT = re.compile(r'^\s*\d+\s+\d+\.\d+\s+(\-?\d+\.?\d*)\s+(\-?\d+\.?\d*)')


class LatLonReader(object):

    def __init__(self, filename):
        self.lonLatTargets = []
        with open(filename) as f:
            for line in f:
                m = PAT.match(line)
                if m:
                    lat, lon = float(m.group(1)), float(m.group(2))
                    self.lonLatTargets.append((lon, lat))

    def getLonLats(self):
        return numpy.array(self.lonLatTargets)


if __name__ == '__main__':
    print(LatLonReader(sys.argv[1]).getLonLats())
