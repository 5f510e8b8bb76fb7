#!/usr/bin/env python3
"""make_vpi_cuda.py -- writes fortran/_gen/vpi_cuda.f90: the reference's own driver `program vpi` (vpi.f90) with its
Monte-Carlo step loop (vpi.f90:297-475) replaced by calls into libpigs_cuda through fortran/pigs_cuda_mod.f90.

The north star keeps the Fortran driver and re-targets it at the C ABI.  The driver is the reference's text, so this
repository does not carry a copy of it: this script reads /root/reference/vpi.f90 where it lies and applies five
edits, each anchored on a line of the reference (the script stops if an anchor is missing):

  1. `use pigs_cuda_mod` + `use iso_c_binding` after the reference's use statements;
  2. declarations of the handle, the parameter / result structs and the scratch arrays (fragment DECLS below);
  3. after the tables are filled (vpi.f90:146-153): create the context, hand over tables and the initial path
     (fragment SETUP);
  4. the whole `do istep=1,Nstep ... end do` (vpi.f90:297-475) becomes ONE call, pigs_run_block, followed by
     pigs_get_block, which returns exactly the sums the loop accumulates (fragment BLOCK);
  5. before CheckPoint (vpi.f90:543) the path of chain 0 is fetched back; Perm_histogram (vpi.f90:590-592) is
     fetched before it is written; the context is destroyed at the end.

Everything else -- ReadParameters, geometry, init, tables, block normalisation, output files, final averages -- stays
the reference's code, compiled from the reference's files.  Build (needs a Fortran compiler, absent in this image):

    python fortran/make_vpi_cuda.py            # -> fortran/_gen/vpi_cuda.f90
    gfortran -O2 -c <reference modules> fortran/pigs_cuda_mod.f90 fortran/_gen/vpi_cuda.f90
    gfortran *.o -L pathintegralgroundstate_b200 -lpigs_cuda -o vpi_cuda_f

The generated program runs one chain with the MT19937 replay stream, i.e. it reproduces `./vpi < vpi.in`; set
PIGS_CHAINS / PIGS_GPUS in the environment for n independent Philox chains on several GPUs.
"""
import os
import re
import sys

REF = os.environ.get("PIGS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

DECLS = """
! ---- libpigs_cuda (added by fortran/make_vpi_cuda.py)
type(c_ptr)             :: gpu
type(pigs_params)       :: gpar
type(pigs_block_result) :: gres
integer(c_int)          :: grc, gopen, gworm, giperm, gnew, gend
integer (kind=4)        :: gchains, ggpus, gch
character (len=32)      :: genv
real (kind=8),dimension(:,:),allocatable  :: gnr
integer (kind=4),dimension(:),allocatable :: gcyc, ghist
"""

SETUP = """
! ---- libpigs_cuda: context, tables, initial state (added by fortran/make_vpi_cuda.py)
gchains = 1
ggpus   = 1
call get_environment_variable('PIGS_CHAINS',genv)
if (len_trim(genv)>0) read (genv,*) gchains
call get_environment_variable('PIGS_GPUS',genv)
if (len_trim(genv)>0) read (genv,*) ggpus

gpar%dim = dim;   gpar%Np = Np;   gpar%Nb = Nb;   gpar%Nmax = Nmax
gpar%Nbin = Nbin; gpar%Nk = Nk;   gpar%Npw = Npw
gpar%trap = merge(1,0,trap)
gpar%Lbox = 0.d0; gpar%a_ho = 1.d0
if (trap) then
   gpar%a_ho(1:dim) = a_ho(1:dim)
else
   gpar%Lbox(1:dim) = Lbox(1:dim)
end if
gpar%rcut = rcut; gpar%dr = dr; gpar%density = density; gpar%dt = dt
gpar%delta_cm = delta_cm; gpar%CWorm = CWorm
gpar%CMFreq = CMFreq
gpar%sampling = merge(0,1,sampling=="sta")
gpar%Lstag = Lstag; gpar%Nlev = Nlev; gpar%Nstag = Nstag; gpar%Nobdm = Nobdm
gpar%swapping = merge(1,0,swapping)
gpar%n_chains = gchains
gpar%rng_mode = merge(PIGS_RNG_MT_REPLAY,PIGS_RNG_PHILOX,gchains==1)
gpar%seed = seed
gpar%device = 0; gpar%threads_per_chain = 0; gpar%table_mode = -1; gpar%action = 0
gpar%schedule = -1; gpar%chain_offset = 0; gpar%gpus = ggpus

grc = pigs_create(gpar,gpu)
if (grc/=0) stop 'pigs_create failed'
grc = pigs_set_tables(gpu,LogWF,VTable)
do gch=0,gchains-1
   ! every chain starts from the configuration `init` built; chain c draws stream seed+c
   grc = pigs_set_state(gpu,gch,Path,xend,merge(1,0,isopen),iworm)
   grc = pigs_sgrnd(gpu,gch,seed+gch)
end do
allocate (gnr(0:Npw,Nbin),gcyc(Np),ghist(Np))
"""

BLOCK = """
   ! ---- the step loop of vpi.f90:297-475, on the GPU (added by fortran/make_vpi_cuda.py)
   grc = pigs_run_block(gpu,Nstep)
   if (grc/=0) stop 'pigs_run_block failed'
   grc = pigs_get_block(gpu,gres,gr,Sk,gnr)

   BlockAvE   = gres%sumE;   BlockAvK   = gres%sumK;   BlockAvV   = gres%sumV
   BlockAvE2  = gres%sumE2;  BlockAvK2  = gres%sumK2;  BlockAvV2  = gres%sumV2
   BlockAvEt  = gres%sumEt;  BlockAvKt  = gres%sumKt;  BlockAvVt  = gres%sumVt
   BlockAvEt2 = gres%sumEt2; BlockAvKt2 = gres%sumKt2; BlockAvVt2 = gres%sumVt2

   idiag_block = int(gres%idiag_block)
   idiag       = idiag+idiag_block
   idiag_aux   = idiag_aux+idiag_block
   ngr         = int(gres%ngr)
   nrho        = nrho+gnr

   try_cm   = real(gres%try_cm,8);        acc_cm   = int(gres%acc_cm)
   try_stag = real(gres%try_stag,8);      acc_bd   = int(gres%acc_bd)
   acc_head = int(gres%acc_head);         acc_tail = int(gres%acc_tail)
   try_cm_half   = real(gres%try_cm_half,8);   acc_cm_half   = int(gres%acc_cm_half)
   try_stag_half = real(gres%try_stag_half,8); acc_bd_half   = int(gres%acc_bd_half)
   acc_head_half = int(gres%acc_head_half);    acc_tail_half = int(gres%acc_tail_half)
   try_open  = int(gres%try_open);  acc_open  = int(gres%acc_open)
   try_close = int(gres%try_close); acc_close = int(gres%acc_close)
   try_swap  = int(gres%try_swap);  acc_swap  = int(gres%acc_swap)
"""

FETCH = """
      ! ---- chain 0 back into the driver's arrays for the reference's CheckPoint (added)
      grc = pigs_get_state(gpu,0,Path,xend,gopen,gworm)
      isopen = gopen/=0
      iworm  = gworm
"""

HIST = """
! ---- Perm_histogram summed over the chains (added by fortran/make_vpi_cuda.py)
if (swapping) then
   Perm_histogram = 0
   do gch=0,gchains-1
      grc = pigs_get_perm(gpu,gch,giperm,gcyc,ghist,gnew,gend)
      Perm_histogram = Perm_histogram+ghist
   end do
end if
"""


def main():
    src = open(os.path.join(REF, "vpi.f90")).read().split("\n")

    def find(pattern, start=0):
        rx = re.compile(pattern)
        for i in range(start, len(src)):
            if rx.search(src[i]):
                return i
        sys.exit(f"make_vpi_cuda: anchor {pattern!r} not found in {REF}/vpi.f90")

    out = list(src)
    # 5c. destroy before the program ends
    i = find(r"^\s*end program vpi")
    out[i:i] = ["grc = pigs_destroy(gpu)", ""]
    # 5b. histogram before it is written
    i = find(r"^\s*do ip=1,Np\s*$", find(r"^\s*end do\s*$", find(r"call cpu_time\(end\)")))
    out[i:i] = HIST.strip("\n").split("\n") + [""]
    # 5a. chain 0 back before CheckPoint
    i = find(r"call CheckPoint\(trap,Path,xend,isopen,iworm\)")
    out[i:i] = FETCH.strip("\n").split("\n") + [""]
    # 4. the step loop
    a = find(r"^\s*do istep=1,Nstep\s*$")
    b = find(r"^\s*if \(idiag_block/=0\) then", a)
    while not re.match(r"^\s*end do\s*$", out[b]):
        b -= 1
    out[a:b + 1] = BLOCK.strip("\n").split("\n")
    # 3. context after the tables
    i = find(r"call PotentialTable\(rcut,VTable\)")
    i = find(r"^\s*end if", i) + 1
    out[i:i] = SETUP.strip("\n").split("\n") + [""]
    # 2. declarations before the first executable statement
    i = find(r"^!Reading input parameters")
    out[i:i] = DECLS.strip("\n").split("\n") + [""]
    # 1. modules
    i = find(r"^use vpi_mod")
    out[i + 1:i + 1] = ["use pigs_cuda_mod", "use, intrinsic :: iso_c_binding"]
    os.makedirs(os.path.join(HERE, "_gen"), exist_ok=True)
    dst = os.path.join(HERE, "_gen", "vpi_cuda.f90")
    open(dst, "w").write("\n".join(out))
    print(f"wrote {dst}: {len(out)} lines ({len(src)} in the reference's vpi.f90)")


if __name__ == "__main__":
    main()
