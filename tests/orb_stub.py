"""TEST INFRASTRUCTURE: a numpy stand-in for one rank's ORB services (PKD.pkdOrbLoad / pkdCalcBound / pkdWeight /
pkdOrbSplit / pkdOrbCells), so that the host-side driver domain.pst_domain_decomp can be exercised on CPU (single
process and across gloo processes).  The product's services are the CUDA kernels of gasoline_b200/csrc/gg_orb.cu;
nothing in gasoline_b200/ imports this."""
import numpy as np

FLOAT_MAXVAL = 1.7976931348623157e308


class HostOrbRank:
    def __init__(self, x, y, z, fWeight=None):
        self.pos = np.stack([np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(z, np.float64)], axis=1)
        self.w = np.ones(len(self.pos)) if fWeight is None else np.asarray(fWeight, np.float64)
        self.cell = np.ones(len(self.pos), np.int32)

    def pkdCalcBound(self, iCell):
        bnd, nIn = np.zeros((len(iCell), 6)), np.zeros(len(iCell), np.int32)
        for k, c in enumerate(iCell):
            p = self.pos[self.cell == c]
            nIn[k] = len(p)
            bnd[k, :3] = p.min(axis=0) if len(p) else FLOAT_MAXVAL
            bnd[k, 3:] = p.max(axis=0) if len(p) else -FLOAT_MAXVAL
        return bnd, nIn

    def pkdWeight(self, iCell, iDim, fSplit):
        k = len(iCell)
        nLow, nHigh, fLow, fHigh = np.zeros(k, np.int32), np.zeros(k, np.int32), np.zeros(k), np.zeros(k)
        for j in range(k):
            m = self.cell == iCell[j]
            low = self.pos[m, iDim[j]] < fSplit[j]
            nLow[j], nHigh[j] = np.count_nonzero(low), np.count_nonzero(~low)
            fLow[j], fHigh[j] = self.w[m][low].sum(), self.w[m][~low].sum()
        return nLow, nHigh, fLow, fHigh

    def pkdOrbSplit(self, iCell, iDim, fSplit):
        for j in range(len(iCell)):
            m = self.cell == iCell[j]
            self.cell[m] = 2 * iCell[j] + (self.pos[m, iDim[j]] >= fSplit[j])

    def pkdOrbSplitWrap(self, iCell, iDim, fSplit, fSplitInactive):
        """The outcome of pkdColRejects with a second boundary (pkd.c:1463-1485): the lower child takes the wrapped interval
        between fSplitInactive and fSplit (pkdLowerPartWrap, pkd.c:1165-1211)."""
        for j in range(len(iCell)):
            m = self.cell == iCell[j]
            c = self.pos[m, iDim[j]]
            if fSplitInactive[j] > fSplit[j]:
                low = (c < fSplit[j]) | (c >= fSplitInactive[j])
            else:
                low = (c < fSplit[j]) & (c >= fSplitInactive[j])
            self.cell[m] = 2 * iCell[j] + (~low)

    def pkdOrbCells(self):
        return self.cell
