"""fluxviz -- drop-in for the COMPUTE STEP of nemoflux/fluxviz.py (same flags).

    python -m nemoflux_b200.fluxviz -t T.nc -u U.nc -v V.nc --lonLatPoints="(-180,-70),(-160,-10),(-35,40)"
    python -m nemoflux_b200.fluxviz -t T.nc -u U.nc -v V.nc -s -i "['S1.txt', 'S2.txt']"

Parses the transects exactly like fluxviz.py:368-397 (plus the flat point list the README uses, README.md:32),
builds the GPU-backed Field and prints what the viewer shows in its title bar: ``flux = <values> (units) <time>``
for time index 0, and for every following index when ``--allTimes`` is given (the 't' key of the viewer,
fluxviz.py:39-44,93-108).  The VTK rendering itself (tubes, glyphs, camera keys) is out of scope: it is not a
data-parallel path and VTK is not part of this image.
"""
import argparse
import ast
import glob

import numpy

from .field import Field, parseLonLatPoints
from .latlonreader import LatLonReader


def parseTransects(lonLatPoints='', iFiles=''):
    """-> (list of (n,3) arrays, list of names)"""
    if lonLatPoints:
        pts = parseLonLatPoints(lonLatPoints)
        return pts, [f'line{i}' for i in range(len(pts))]
    if iFiles:
        try:
            listOfFiles = ast.literal_eval(iFiles)
            if isinstance(listOfFiles, str):
                listOfFiles = [listOfFiles]
        except (ValueError, SyntaxError):
            listOfFiles = sorted(glob.glob(iFiles)) or [iFiles]
        print(f'list of target surfaces: {listOfFiles}')
        pts = []
        for iFile in listOfFiles:
            pts.append(numpy.array([(ll[0], ll[1], 0.) for ll in LatLonReader(iFile).getLonLats()], numpy.float64))
        return pts, list(listOfFiles)
    raise RuntimeError('ERROR must provide either iFiles (-i) or lonLatPoints (-l)!')


class FluxViz(object):
    """compute-only counterpart of fluxviz.FluxViz: holds the Field, steps in time, produces the title text"""

    def __init__(self, tFile, uFile, vFile, lonLatZPoints, sverdrup=False, meshFile=None):
        self.field = Field(tFile, uFile, vFile, lonLatZPoints, sverdrup, meshFile=meshFile)

    def update(self, key):
        if key == 't':
            self.field.timeIndex = (self.field.timeIndex + 1) % self.field.nt
        elif key == 'T':
            self.field.timeIndex = (self.field.timeIndex - 1) % self.field.nt
        else:
            return
        self.field.update()
        nvpts = self.field.vectorPoints.shape[0]
        if nvpts:      # arrows are drawn normalised by the longest one (fluxviz.py:95-97)
            maxVectorLength = numpy.sqrt((self.field.vectorValues ** 2).sum(axis=1)).max()
            if maxVectorLength > 0:
                self.field.vectorValues /= maxVectorLength
        print(f'time index now {self.field.timeIndex} max |flux|: {self.field.maxAbsFlux:10.3f} nt = {self.field.nt}')

    def title(self):
        f = self.field
        return f'flux = {f.getFluxText()} {f.timeObj.getTimeAsString(f.timeIndex)}'

    def show(self, allTimes=False):
        print(self.title())
        if allTimes:
            for _ in range(self.field.nt - 1):
                self.update('t')
                print(self.title())


def main(argv=None):
    ap = argparse.ArgumentParser(description='Visualize fluxes (compute step)')
    ap.add_argument('-t', '--tFile', required=True, help='netcdf file holding the T-grid')
    ap.add_argument('-u', '--uFile', required=True, help='netcdf file holding u data')
    ap.add_argument('-v', '--vFile', required=True, help='netcdf file holding v data')
    ap.add_argument('-l', '--lonLatPoints', default='', help='target points "[(lon0, lat0), (lon1, lat1),...],[...]"')
    ap.add_argument('-i', '--iFiles', default='', help='alternatively read target points from text files')
    ap.add_argument('-s', '--sverdrup', action='store_true', help='use Sverdrup units (default is A m^2/s)')
    ap.add_argument('--allTimes', action='store_true', help='step through every time index')
    ap.add_argument('--meshFile', default='', help='optional mesh_mask file: e3u_0/e3v_0 (and e2u/e1v) scale factors')
    a = ap.parse_args(argv)
    pts, _ = parseTransects(a.lonLatPoints, a.iFiles)
    print(f'target points:\n {pts}')
    fv = FluxViz(a.tFile, a.uFile, a.vFile, pts, a.sverdrup, meshFile=a.meshFile or None)
    fv.show(a.allTimes)
    return fv


if __name__ == '__main__':
    main()
