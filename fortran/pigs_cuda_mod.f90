!-----------------------------------------------------------------------
! pigs_cuda_mod -- ISO_C_BINDING interface of libpigs_cuda (include/pigs_cuda.h)
! for the reference driver vpi.f90.
!
! NOT compiled in this repository's image (no Fortran compiler exists here:
! gfortran/flang/ifort/nvfortran are all absent), so the file is kept purely
! mechanical: one bind(C) interface per C entry point, derived types that
! mirror the C structs field for field.  Build with the reference's makefile:
!     gfortran -c pigs_cuda_mod.f90   and link   -L... -lpigs_cuda
! Logicals cross the boundary as integer(c_int) (gfortran's default logical is
! 4 bytes but its bit pattern is not part of any standard).
!-----------------------------------------------------------------------
module pigs_cuda_mod

use, intrinsic :: iso_c_binding

implicit none

integer(c_int), parameter :: PIGS_RNG_PHILOX = 0, PIGS_RNG_MT_REPLAY = 1

integer(c_int), parameter :: PIGS_TRANSLATE_CHAIN = 0, PIGS_STAGING = 1, PIGS_MOVE_HEAD = 2, &
     PIGS_MOVE_TAIL = 3, PIGS_BISECTION = 4, PIGS_MOVE_HEAD_BISECTION = 5, PIGS_MOVE_TAIL_BISECTION = 6, &
     PIGS_TRANSLATE_HALF = 7, PIGS_STAGING_HALF = 8, PIGS_MOVE_HEAD_HALF = 9, PIGS_MOVE_TAIL_HALF = 10, &
     PIGS_OPEN = 11, PIGS_CLOSE = 12, PIGS_SWAP = 13

! struct pigs_params
type, bind(C) :: pigs_params
   integer(c_int32_t) :: dim, Np, Nb, Nmax, Nbin, Nk, Npw, trap
   real(c_double)     :: Lbox(3), a_ho(3)
   real(c_double)     :: rcut, dr, density, dt, delta_cm, CWorm
   integer(c_int32_t) :: CMFreq, sampling, Lstag, Nlev, Nstag, Nobdm, swapping
   integer(c_int32_t) :: n_chains, rng_mode
   integer(c_int64_t) :: seed
   integer(c_int32_t) :: device, threads_per_chain, table_mode, action
   integer(c_int32_t) :: schedule, chain_offset, gpus      ! -1 / 0 / 1 for the reference's single-GPU, single-run use
end type pigs_params

! struct pigs_block_result
type, bind(C) :: pigs_block_result
   real(c_double)     :: sumE, sumK, sumV, sumEt, sumKt, sumVt
   real(c_double)     :: sumE2, sumK2, sumV2, sumEt2, sumKt2, sumVt2
   integer(c_int64_t) :: idiag_block, ngr
   integer(c_int64_t) :: try_cm, try_stag, try_cm_half, try_stag_half
   integer(c_int64_t) :: acc_cm, acc_bd, acc_head, acc_tail
   integer(c_int64_t) :: acc_cm_half, acc_bd_half, acc_head_half, acc_tail_half
   integer(c_int64_t) :: try_open, acc_open, try_close, acc_close, try_swap, acc_swap
   integer(c_int64_t) :: bead_updates(3)
   integer(c_int64_t) :: n_open_chains
end type pigs_block_result

interface

   function pigs_create(p, h) bind(C, name='pigs_create') result(rc)
     import :: c_int, c_ptr, pigs_params
     type(pigs_params), intent(in) :: p
     type(c_ptr), intent(out)      :: h
     integer(c_int)                :: rc
   end function pigs_create

   function pigs_destroy(h) bind(C, name='pigs_destroy') result(rc)
     import :: c_int, c_ptr
     type(c_ptr), value :: h
     integer(c_int)     :: rc
   end function pigs_destroy

   function pigs_last_error() bind(C, name='pigs_last_error') result(msg)
     import :: c_ptr
     type(c_ptr) :: msg            ! NUL-terminated C string
   end function pigs_last_error

   ! JastrowTable / PotentialTable results: LogWF(0:Nmax+1), VTable(0:Nmax+1)
   function pigs_set_tables(h, LogWF, VTable) bind(C, name='pigs_set_tables') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value         :: h
     real(c_double), intent(in) :: LogWF(*), VTable(*)
     integer(c_int)             :: rc
   end function pigs_set_tables

   ! Path(dim,Np,0:2*Nb), xend(dim,2) exactly as the driver holds them
   function pigs_set_state(h, chain, Path, xend, isopen, iworm) bind(C, name='pigs_set_state') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value         :: h
     integer(c_int), value      :: chain, isopen, iworm
     real(c_double), intent(in) :: Path(*), xend(*)
     integer(c_int)             :: rc
   end function pigs_set_state

   function pigs_get_state(h, chain, Path, xend, isopen, iworm) bind(C, name='pigs_get_state') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value          :: h
     integer(c_int), value       :: chain
     real(c_double), intent(out) :: Path(*), xend(*)
     integer(c_int), intent(out) :: isopen, iworm
     integer(c_int)              :: rc
   end function pigs_get_state

   function pigs_sgrnd(h, chain, seed) bind(C, name='pigs_sgrnd') result(rc)
     import :: c_int, c_ptr, c_int32_t
     type(c_ptr), value        :: h
     integer(c_int), value     :: chain
     integer(c_int32_t), value :: seed
     integer(c_int)            :: rc
   end function pigs_sgrnd

   ! the step loop of vpi.f90:297-475 for Nstep steps on every chain
   function pigs_run_block(h, Nstep) bind(C, name='pigs_run_block') result(rc)
     import :: c_int, c_ptr
     type(c_ptr), value    :: h
     integer(c_int), value :: Nstep
     integer(c_int)        :: rc
   end function pigs_run_block

   ! gr(Nbin), Sk(dim,Nk), nrho(0:Npw,Nbin): raw sums of the block over all chains
   function pigs_get_block(h, res, gr, Sk, nrho) bind(C, name='pigs_get_block') result(rc)
     import :: c_int, c_ptr, c_double, pigs_block_result
     type(c_ptr), value                   :: h
     type(pigs_block_result), intent(out) :: res
     real(c_double), intent(out)          :: gr(*), Sk(*), nrho(*)
     integer(c_int)                       :: rc
   end function pigs_get_block

   function pigs_get_perm(h, chain, iperm, cycle, hist, new_pc, end_pc) bind(C, name='pigs_get_perm') result(rc)
     import :: c_int, c_ptr, c_int32_t
     type(c_ptr), value              :: h
     integer(c_int), value           :: chain
     integer(c_int), intent(out)     :: iperm, new_pc, end_pc
     integer(c_int32_t), intent(out) :: cycle(*), hist(*)
     integer(c_int)                  :: rc
   end function pigs_get_perm

   ! one reference procedure on every chain (unit API)
   function pigs_move(h, move, ip, half, accepted, aux) bind(C, name='pigs_move') result(rc)
     import :: c_int, c_ptr, c_int32_t
     type(c_ptr), value              :: h
     integer(c_int), value           :: move, ip, half
     integer(c_int32_t), intent(out) :: accepted(*), aux(*)
     integer(c_int)                  :: rc
   end function pigs_move

   ! UpdateAction on caller data: R(dim,Np,n), ip(n), ib(n), xnew(dim,n), xold(dim,n) -> DeltaS(n)
   function pigs_update_action(h, n, R, ip, ib, xnew, xold, DeltaS) bind(C, name='pigs_update_action') result(rc)
     import :: c_int, c_ptr, c_double, c_int32_t
     type(c_ptr), value             :: h
     integer(c_int), value          :: n
     real(c_double), intent(in)     :: R(*), xnew(*), xold(*)
     integer(c_int32_t), intent(in) :: ip(*), ib(*)
     real(c_double), intent(out)    :: DeltaS(*)
     integer(c_int)                 :: rc
   end function pigs_update_action

   ! LocalEnergy on caller data: R(dim,Np,n) -> E(n),Kin(n),Pot(n)
   function pigs_local_energy(h, n, R, E, Kin, Pot) bind(C, name='pigs_local_energy') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value          :: h
     integer(c_int), value       :: n
     real(c_double), intent(in)  :: R(*)
     real(c_double), intent(out) :: E(*), Kin(*), Pot(*)
     integer(c_int)              :: rc
   end function pigs_local_energy

   ! ThermEnergy on caller data: Path(dim,Np,0:2*Nb,n) -> E(n),Ec(n),Ep(n)
   function pigs_therm_energy(h, n, Path, E, Ec, Ep) bind(C, name='pigs_therm_energy') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value          :: h
     integer(c_int), value       :: n
     real(c_double), intent(in)  :: Path(*)
     real(c_double), intent(out) :: E(*), Ec(*), Ep(*)
     integer(c_int)              :: rc
   end function pigs_therm_energy

   function pigs_pair_correlation(h, n, R, gr) bind(C, name='pigs_pair_correlation') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value            :: h
     integer(c_int), value         :: n
     real(c_double), intent(in)    :: R(*)
     real(c_double), intent(inout) :: gr(*)
     integer(c_int)                :: rc
   end function pigs_pair_correlation

   function pigs_structure_factor(h, n, R, Sk) bind(C, name='pigs_structure_factor') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value            :: h
     integer(c_int), value         :: n
     real(c_double), intent(in)    :: R(*)
     real(c_double), intent(inout) :: Sk(*)
     integer(c_int)                :: rc
   end function pigs_structure_factor

   function pigs_obdm(h, n, xend, nrho) bind(C, name='pigs_obdm') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value            :: h
     integer(c_int), value         :: n
     real(c_double), intent(in)    :: xend(*)
     real(c_double), intent(inout) :: nrho(*)
     integer(c_int)                :: rc
   end function pigs_obdm

   ! MT19937 state of one chain (mtsavef / mtgetf, random_mod.f90:125-191)
   function pigs_get_mt(h, chain, mt624, mti) bind(C, name='pigs_get_mt') result(rc)
     import :: c_int, c_ptr, c_int32_t
     type(c_ptr), value              :: h
     integer(c_int), value           :: chain
     integer(c_int32_t), intent(out) :: mt624(*), mti
     integer(c_int)                  :: rc
   end function pigs_get_mt

   function pigs_set_mt(h, chain, mt624, mti) bind(C, name='pigs_set_mt') result(rc)
     import :: c_int, c_ptr, c_int32_t
     type(c_ptr), value             :: h
     integer(c_int), value          :: chain
     integer(c_int32_t), intent(in) :: mt624(*)
     integer(c_int32_t), value      :: mti
     integer(c_int)                 :: rc
   end function pigs_set_mt

   ! n draws of the chain's grnd() stream (random_mod.f90:35-115), e.g. for init's starting positions
   function pigs_grnd(h, chain, n, u) bind(C, name='pigs_grnd') result(rc)
     import :: c_int, c_ptr, c_double
     type(c_ptr), value          :: h
     integer(c_int), value       :: chain, n
     real(c_double), intent(out) :: u(*)
     integer(c_int)              :: rc
   end function pigs_grnd

   ! every chain's complete state in one binary file (what CheckPoint + mtsavef do for one chain)
   function pigs_save_checkpoint(h, path) bind(C, name='pigs_save_checkpoint') result(rc)
     import :: c_int, c_ptr, c_char
     type(c_ptr), value                 :: h
     character(kind=c_char), intent(in) :: path(*)      ! NUL-terminated
     integer(c_int)                     :: rc
   end function pigs_save_checkpoint

   function pigs_load_checkpoint(h, path) bind(C, name='pigs_load_checkpoint') result(rc)
     import :: c_int, c_ptr, c_char
     type(c_ptr), value                 :: h
     character(kind=c_char), intent(in) :: path(*)
     integer(c_int)                     :: rc
   end function pigs_load_checkpoint

end interface

end module pigs_cuda_mod
