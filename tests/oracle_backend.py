"""Adapter giving the CPU oracle the method names of PigsCuda, so that the SAME driver
code (pathintegralgroundstate_b200/driver.py) can be run over the checker and over
the product and the resulting files compared.  Test infrastructure only."""
import numpy as np

from oracle.pigs_oracle import Oracle
from tests.common import oracle_cfg


class OracleBackend:
    n_chains = 1

    def __init__(self, cfg):
        self.o = Oracle(oracle_cfg(cfg))
        self.nr = None

    def set_tables(self, W, V):
        self.o.set_tables(W, V)

    def set_state(self, chain, Path, xend, isopen=0, iworm=0):
        self.o.set_state(Path, xend, isopen, iworm)

    def get_state(self, chain):
        return self.o.get_state()

    def sgrnd(self, seed, chain=0):
        self.o.sgrnd(seed)

    def grnd(self, n, chain=0):
        return np.array([self.o.grnd() for _ in range(n)])

    def get_mt(self, chain):
        return self.o.get_mt()

    def set_mt(self, chain, mt, mti):
        self.o.set_mt(mt, mti)

    def get_perm(self, chain):
        return self.o.get_perm()

    def run_block(self, nstep):
        self._last = self.o.run_block(nstep)

    def get_block(self, chain=None):
        b, gr, Sk, nr = self._last
        b = dict(b)
        b["n_open_chains"] = self.o.get_state()[2]
        return b, gr, Sk, nr
