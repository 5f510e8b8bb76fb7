"""pigs-b200: the PIGS worldline-update + estimator hot path of
amaciarey/PathIntegralGroundState on NVIDIA B200 (sm_100a), behind a C ABI
(include/pigs_cuda.h, libpigs_cuda.so).  See DESIGN.md."""
from .host import (PigsCuda, PigsError, PigsParams, PigsBlockResult, load_library, derive_geometry, make_table,
                   aziz_hfdb, aziz_hfdhe2, lennard_jones, mcmillan_logpsi, write_config_ini, read_config_ini, normalize_gr, normalize_sk, normalize_nr,
                   measure_fp64_peak, MOVES, LIB_PATH)
from .vpi_in import read_vpi_in, parse_namelists

__all__ = ["PigsCuda", "PigsError", "PigsParams", "PigsBlockResult", "load_library", "derive_geometry", "make_table",
           "aziz_hfdb", "aziz_hfdhe2", "lennard_jones", "mcmillan_logpsi", "write_config_ini", "read_config_ini", "normalize_gr", "normalize_sk", "normalize_nr",
           "measure_fp64_peak", "MOVES", "LIB_PATH", "read_vpi_in", "parse_namelists"]
