! afesp_gpu.f90 -- ISO_C_BINDING shim between AFESP's Fortran host (els.x) and libafesp_gpu.so (include/afesp_gpu.h).
!
! Drop this file into src/ of the reference, add it to the CMake source list, build with -DAFESP_CUDA and link
! -lafesp_gpu.  It replaces the bodies of do_mp2_spatial / do_ccsd_spatial / do_ccsd_spinorb / do_ccsd_t_* by calls
! into the CUDA library while keeping els.in, the *.dat readers, the RHF, every printed line and the convergence logic
! in Fortran.  The edited call sites in main.F90 are listed in INTEGRATION.md; they mirror the existing
! `#ifdef OPENACC` switch at main.F90:74-80.
!
! NOTE: this image has no Fortran compiler (SURVEY.md K5), so this file is delivered as source only; the identical ABI
! is exercised from Python/ctypes (afesp_b200/capi.py, tests/) and documented in include/afesp_gpu.h.
module afesp_gpu
   use, intrinsic :: iso_c_binding
   use const, only: p
   use error_handling, only: error
   implicit none
   private
   public :: gpu_open, gpu_close, gpu_mp2, gpu_ccsd, gpu_ccsd_t_spatial, gpu_ccsd_t_spinorb, gpu_handle
   public :: gpu_comm_id, gpu_comm_attach

   type(c_ptr), save :: gpu_handle = c_null_ptr

   interface
      integer(c_int) function afesp_gpu_open(device, h) bind(C, name='afesp_gpu_open')
         import :: c_int, c_ptr
         integer(c_int), value :: device
         type(c_ptr), intent(out) :: h
      end function
      integer(c_int) function afesp_gpu_close(h) bind(C, name='afesp_gpu_close')
         import :: c_int, c_ptr
         type(c_ptr), value :: h
      end function
      type(c_ptr) function afesp_gpu_last_error(h) bind(C, name='afesp_gpu_last_error')
         import :: c_ptr
         type(c_ptr), value :: h
      end function
      integer(c_int) function afesp_gpu_set_option(h, key, val) bind(C, name='afesp_gpu_set_option')
         import :: c_int, c_ptr, c_char, c_double
         type(c_ptr), value :: h
         character(kind=c_char), dimension(*), intent(in) :: key
         real(c_double), value :: val
      end function
      integer(c_int) function afesp_gpu_ao2mo(h, nbasis, eri_ao, coeff, eri_mo) bind(C, name='afesp_gpu_ao2mo')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nbasis
         real(c_double), dimension(*), intent(in) :: eri_ao, coeff
         real(c_double), dimension(*), intent(out) :: eri_mo
      end function
      integer(c_int) function afesp_gpu_mp2_energy(h, nocc, eps, e_mp2) bind(C, name='afesp_gpu_mp2_energy')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nocc
         real(c_double), dimension(*), intent(in) :: eps
         real(c_double), intent(out) :: e_mp2
      end function
      integer(c_int) function afesp_gpu_ccsd_init_info(h, info) bind(C, name='afesp_gpu_ccsd_init_info')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), dimension(4), intent(out) :: info
      end function
      integer(c_int) function afesp_gpu_ccsd_init(h, nocc, restricted, eps, diis_n, e_mp1, rmst2) &
            bind(C, name='afesp_gpu_ccsd_init')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nocc, restricted, diis_n
         real(c_double), dimension(*), intent(in) :: eps
         real(c_double), intent(out) :: e_mp1, rmst2
      end function
      integer(c_int) function afesp_gpu_ccsd_iterate(h, e_cc, rmst2) bind(C, name='afesp_gpu_ccsd_iterate')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: e_cc, rmst2
      end function
      integer(c_int) function afesp_gpu_ccsd_diis(h) bind(C, name='afesp_gpu_ccsd_diis')
         import :: c_int, c_ptr
         type(c_ptr), value :: h
      end function
      integer(c_int) function afesp_gpu_ccsd_finalize(h, want_cr, t1_diag, t1, t2) bind(C, name='afesp_gpu_ccsd_finalize')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: want_cr
         real(c_double), intent(out) :: t1_diag
         type(c_ptr), value :: t1, t2      ! c_null_ptr: keep the amplitudes on the device only
      end function
      integer(c_int) function afesp_gpu_ccsd_t_spatial(h, paren, renorm, comp_renorm, sums, dconst) &
            bind(C, name='afesp_gpu_ccsd_t_spatial')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: paren, renorm, comp_renorm
         real(c_double), dimension(6), intent(out) :: sums
         real(c_double), intent(out) :: dconst
      end function
      integer(c_int) function afesp_gpu_ccsd_t_spinorb(h, e_T) bind(C, name='afesp_gpu_ccsd_t_spinorb')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: e_T
      end function
      ! Multi-GPU (one process per GPU, e.g. one MPI rank each): rank 0 obtains the 128-byte id, the host program
      ! broadcasts it with its own transport (MPI_Bcast), every rank attaches.  Afterwards the stage calls above are
      ! collective: AO->MO, the heavy CCSD GEMMs with the ladder integrals and the (T) triples are sharded over the ranks.
      integer(c_int) function afesp_gpu_comm_unique_id(id) bind(C, name='afesp_gpu_comm_unique_id')
         import :: c_int, c_char
         character(kind=c_char), dimension(128), intent(out) :: id
      end function
      integer(c_int) function afesp_gpu_comm_init(h, rank, nranks, id) bind(C, name='afesp_gpu_comm_init')
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: h
         integer(c_int), value :: rank, nranks
         character(kind=c_char), dimension(128), intent(in) :: id
      end function
      ! ---- the remaining exports of include/afesp_gpu.h (a host that keeps more of the work, measurement, multi-GPU without
      !      NCCL); tests/test_shim_interfaces.py checks every interface of this block against the C prototypes
      integer(c_int) function afesp_gpu_tma_status(h, scope, selftest) bind(C, name='afesp_gpu_tma_status')
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int), intent(out) :: scope, selftest
      end function
      integer(c_int) function afesp_gpu_counters(h, launches, gemm_flops) bind(C, name='afesp_gpu_counters')
         import :: c_int, c_ptr, c_long_long, c_double
         type(c_ptr), value :: h
         integer(c_long_long), intent(out) :: launches
         real(c_double), intent(out) :: gemm_flops
      end function
      integer(c_int) function afesp_gpu_synth_eri_ao(h, nbasis, naux, factors, coeff) bind(C, name='afesp_gpu_synth_eri_ao')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nbasis, naux
         real(c_double), dimension(*), intent(in) :: factors, coeff
      end function
      integer(c_int) function afesp_gpu_get_eri_mo(h, eri_mo) bind(C, name='afesp_gpu_get_eri_mo')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), dimension(*), intent(out) :: eri_mo
      end function
      integer(c_int) function afesp_gpu_release(h, what) bind(C, name='afesp_gpu_release')
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: h
         character(kind=c_char), dimension(*), intent(in) :: what
      end function
      integer(c_int) function afesp_gpu_set_eri_mo(h, nbasis, eri_mo) bind(C, name='afesp_gpu_set_eri_mo')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nbasis
         real(c_double), dimension(*), intent(in) :: eri_mo
      end function
      integer(c_int) function afesp_gpu_host_register(ptr, bytes) bind(C, name='afesp_gpu_host_register')
         import :: c_int, c_ptr, c_long_long
         type(c_ptr), value :: ptr          ! c_loc(int_store%eri_mo)
         integer(c_long_long), value :: bytes
      end function
      integer(c_int) function afesp_gpu_host_unregister(ptr) bind(C, name='afesp_gpu_host_unregister')
         import :: c_int, c_ptr
         type(c_ptr), value :: ptr
      end function
      integer(c_int) function afesp_gpu_set_partition(h, rank, nranks) bind(C, name='afesp_gpu_set_partition')
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int), value :: rank, nranks
      end function
      integer(c_int) function afesp_gpu_triples_partition(nocc_active, symmetric, strict, nranks, counts) &
            bind(C, name='afesp_gpu_triples_partition')
         import :: c_int, c_long_long
         integer(c_int), value :: nocc_active, symmetric, strict, nranks
         integer(c_long_long), dimension(*), intent(out) :: counts
      end function
      integer(c_int) function afesp_gpu_column_partition(ncols, nranks, granularity, lo, hi) &
            bind(C, name='afesp_gpu_column_partition')
         import :: c_int, c_long_long
         integer(c_long_long), value :: ncols
         integer(c_int), value :: nranks, granularity
         integer(c_long_long), dimension(*), intent(out) :: lo, hi
      end function
      ! the operators of src/linalg.fpp on host arrays (dgemm_wrapper :58-89, omp_reshape :99-156)
      integer(c_int) function afesp_gpu_dgemm_wrapper(h, transA, transB, outer_row, outer_col, inner_dim, A, B, C, alpha, beta) &
            bind(C, name='afesp_gpu_dgemm_wrapper')
         import :: c_int, c_ptr, c_char, c_double
         type(c_ptr), value :: h
         character(kind=c_char), value :: transA, transB
         integer(c_int), value :: outer_row, outer_col, inner_dim
         real(c_double), dimension(*), intent(in) :: A, B
         real(c_double), dimension(*), intent(inout) :: C
         real(c_double), value :: alpha, beta
      end function
      integer(c_int) function afesp_gpu_omp_reshape(h, out_arr, in_arr, in_dims, arr_order, has_beta, beta) &
            bind(C, name='afesp_gpu_omp_reshape')
         import :: c_int, c_ptr, c_char, c_double
         type(c_ptr), value :: h
         real(c_double), dimension(*), intent(inout) :: out_arr
         real(c_double), dimension(*), intent(in) :: in_arr
         integer(c_int), dimension(4), intent(in) :: in_dims
         character(kind=c_char), dimension(4), intent(in) :: arr_order
         integer(c_int), value :: has_beta
         real(c_double), value :: beta
      end function
      ! measurement aids (bench.py uses them through ctypes)
      integer(c_int) function afesp_gpu_bench_dgemm(h, transA, transB, M, N, K, beta, reps, ms) &
            bind(C, name='afesp_gpu_bench_dgemm')
         import :: c_int, c_ptr, c_char, c_double
         type(c_ptr), value :: h
         character(kind=c_char), value :: transA, transB
         integer(c_int), value :: M, N, K
         real(c_double), value :: beta
         integer(c_int), value :: reps
         real(c_double), intent(out) :: ms
      end function
      integer(c_int) function afesp_gpu_bench_hbm(h, what, nocc, nvirt, reps, ms, bytes) bind(C, name='afesp_gpu_bench_hbm')
         import :: c_int, c_ptr, c_char, c_double
         type(c_ptr), value :: h
         character(kind=c_char), dimension(*), intent(in) :: what
         integer(c_int), value :: nocc, nvirt, reps
         real(c_double), intent(out) :: ms, bytes
      end function
      integer(c_int) function afesp_gpu_gemm_crosscheck(h, transA, transB, M, N, K, nbatch, beta, reps, mismatches, ms_tma, &
            ms_cpasync) bind(C, name='afesp_gpu_gemm_crosscheck')
         import :: c_int, c_ptr, c_char, c_double, c_long_long
         type(c_ptr), value :: h
         character(kind=c_char), value :: transA, transB
         integer(c_int), value :: M, N, K, nbatch
         real(c_double), value :: beta
         integer(c_int), value :: reps
         integer(c_long_long), intent(out) :: mismatches
         real(c_double), intent(out) :: ms_tma, ms_cpasync
      end function
      integer(c_int) function afesp_gpu_dmma_peak(h, tflops) bind(C, name='afesp_gpu_dmma_peak')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: tflops
      end function
      integer(c_int) function afesp_gpu_last_stage_ms(h, ms) bind(C, name='afesp_gpu_last_stage_ms')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: ms
      end function
      integer(c_int) function afesp_gpu_timer(h, stop, ms) bind(C, name='afesp_gpu_timer')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: stop
         real(c_double), intent(out) :: ms
      end function
      integer(c_int) function afesp_gpu_gemm_time(h, ms, flops) bind(C, name='afesp_gpu_gemm_time')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: ms, flops
      end function
      integer(c_int) function afesp_gpu_gemm_stats(h, ms, flops, launches) bind(C, name='afesp_gpu_gemm_stats')
         import :: c_int, c_ptr, c_double, c_long_long
         type(c_ptr), value :: h
         real(c_double), intent(out) :: ms, flops
         integer(c_long_long), intent(out) :: launches
      end function
   end interface

contains

   subroutine check(status, where)
      ! Non-zero status -> the reference's own error path: message on stderr, stop 999 (src/error_handling.f90:7-20)
      integer(c_int), intent(in) :: status
      character(*), intent(in) :: where
      character(kind=c_char), pointer :: msg(:)
      character(512) :: text
      integer :: i
      if (status == 0) return
      text = ''
      call c_f_pointer(afesp_gpu_last_error(gpu_handle), msg, [512])
      do i = 1, 512
         if (msg(i) == c_null_char) exit
         text(i:i) = msg(i)
      end do
      call error('afesp_gpu::'//where, trim(text))
   end subroutine check

   subroutine gpu_open(device)
      integer, intent(in) :: device
      call check(afesp_gpu_open(int(device, c_int), gpu_handle), 'open')
   end subroutine gpu_open

   subroutine gpu_comm_id(id)
      ! rank 0 only; broadcast `id` to the other ranks before gpu_comm_attach
      character(kind=c_char), dimension(128), intent(out) :: id
      call check(afesp_gpu_comm_unique_id(id), 'comm_unique_id')
   end subroutine gpu_comm_id

   subroutine gpu_comm_attach(rank, nranks, id)
      integer, intent(in) :: rank, nranks
      character(kind=c_char), dimension(128), intent(in) :: id
      call check(afesp_gpu_comm_init(gpu_handle, int(rank, c_int), int(nranks, c_int), id), 'comm_init')
   end subroutine gpu_comm_attach

   subroutine gpu_close()
      call check(afesp_gpu_close(gpu_handle), 'close')
      gpu_handle = c_null_ptr
   end subroutine gpu_close

   subroutine gpu_mp2(sys, int_store)
      ! Replaces the body of do_mp2_spatial (src/mp2.f90:261-449): AO->MO transform + MP2 energy, same printed lines.
      use, intrinsic :: iso_fortran_env, only: iunit => output_unit
      use system, only: system_t
      use integrals, only: int_store_t
      type(system_t), intent(inout) :: sys
      type(int_store_t), intent(inout) :: int_store
      real(c_double) :: emp
      write(iunit, '(1X, 10("-"))'); write(iunit, '(1X, A)') 'MP2'; write(iunit, '(1X, 10("-"))')
      write(iunit, '(1X, A)') 'Performing AO to MO ERI transformation...'
      allocate(int_store%eri_mo, mold=int_store%eri)
      call check(afesp_gpu_ao2mo(gpu_handle, int(sys%nbasis, c_int), int_store%eri, sys%canon_coeff, int_store%eri_mo), 'ao2mo')
      write(iunit, '(1X, A)') 'Calculating MP2 energy...'
      call check(afesp_gpu_mp2_energy(gpu_handle, int(sys%nel/2, c_int), sys%canon_levels, emp), 'mp2_energy')
      sys%e_mp2 = emp
      sys%e_highest = emp
      write(iunit, '(1X, A, 1X, F15.8)') 'MP2 correlation energy (Hartree):', sys%e_mp2
   end subroutine gpu_mp2

   subroutine gpu_ccsd(sys, restricted)
      ! Replaces do_ccsd_spatial (src/ccsd.f90:279-402) and do_ccsd_spinorb (:71-277): the iteration table, the
      ! convergence test (:1805) and the DIIS call order stay here, one GPU call per step.
      use, intrinsic :: iso_fortran_env, only: iunit => output_unit
      use system, only: system_t
      type(system_t), intent(inout) :: sys
      logical, intent(in) :: restricted
      real(c_double) :: e, e_old, rms, t1d, info(4)
      integer :: iter
      integer(c_int) :: rc
      integer(kind=8) :: t0, t1, c_rate
      logical :: conv
      write(iunit, '(1X, 10("-"))'); write(iunit, '(1X, A)') 'CCSD'; write(iunit, '(1X, 10("-"))')
      call system_clock(count=t0, count_rate=c_rate)
      rc = afesp_gpu_ccsd_init(gpu_handle, int(sys%nel/2, c_int), merge(1_c_int, 0_c_int, restricted), &
                               sys%canon_levels, int(sys%ccsd_diis_n_errmat, c_int), e, rms)
      if (.not. restricted) then
         ! src/ccsd.f90:106-202: the slices of <pq||rs> are gathered straight from the packed MO integrals and the
         ! reference's permutational-symmetry assertion (:150-167) runs on the device; status 5 = it fired
         if (rc == 0 .or. rc == 5) call check(afesp_gpu_ccsd_init_info(gpu_handle, info), 'ccsd_init_info')
         write(iunit, '(1X, A)') 'Forming antisymmetrised spinorbital ERIs...'
         write(iunit, '(1X, A, 1X, F8.6, A)') 'Time taken:', info(2), " s"
         write(iunit, *)
         write(iunit, '(1X, A)') 'Checking that the permuational symmetry of the antisymmetrised integrals hold...'
         if (rc == 5) then
            write(iunit, '(1X, A, 1X, E15.6)') 'Permutational symmetry error:', info(1)
            call error('ccsd::do_ccsd', 'Permutational symmetry of antisymmetrised integrals does not hold')
         end if
         write(iunit, '(1X, A, 1X, F8.6, A)') 'Time taken:', info(3), " s"
         write(iunit, *)
      end if
      call check(rc, 'ccsd_init')
      write(iunit, '(75("-"))')
      write(iunit, '(1X, A, 3X, A, 3X, A, 3X, A, 3X, A)') &
         'Iteration','     Energy    ','    deltaE     ','  delta RMS T2 ', '  Time  '
      write(iunit, '(75("-"))')
      write(iunit, '(1X, A9, 3X, F15.12, 3X, F15.12, 3X, F15.12)') 'MP1', e, e, rms
      conv = .false.
      do iter = 1, sys%ccsd_maxiter
         e_old = e
         call check(afesp_gpu_ccsd_iterate(gpu_handle, e, rms), 'ccsd_iterate')
         call system_clock(t1)
         write(iunit, '(1X, I9, 3X, F15.12, 3X, F15.12, 3X, F15.12, 3X, F8.6)') iter, e, e-e_old, rms, real(t1-t0, kind=p)/c_rate
         t0 = t1
         if (sqrt(rms) < sys%ccsd_t_tol .and. abs(e-e_old) < sys%ccsd_e_tol) then
            conv = .true.
            write(iunit, '(75("-"))')
            write(iunit, '(1X, A)') 'Convergence reached within tolerance.'
            write(iunit, '(1X, A, 1X, F15.12)') 'Final CCSD Energy (Hartree):', e
            call check(afesp_gpu_ccsd_finalize(gpu_handle, merge(1_c_int, 0_c_int, sys%ccsd_t_comp_renorm), t1d, &
                                               c_null_ptr, c_null_ptr), 'ccsd_finalize')
            if (restricted) then
               sys%t1_diagnostic = t1d
               write(iunit, '(1X, A, 1X, F8.5)') 'T1 diagnostic:', sys%t1_diagnostic
               if (sys%t1_diagnostic > 0.02_p) write(iunit, '(1X, A)') &
                  'Significant multireference character detected, CCSD result might be unreliable!'
            end if
            sys%e_ccsd = e
            sys%e_highest = e
            exit
         end if
         call check(afesp_gpu_ccsd_diis(gpu_handle), 'ccsd_diis')
      end do
   end subroutine gpu_ccsd

   subroutine gpu_ccsd_t_spatial(sys, calcname)
      ! Replaces do_ccsd_t_spatial (src/ccsd.f90:2018-2293); the energy assembly is :2239-2276 verbatim.
      use, intrinsic :: iso_fortran_env, only: iunit => output_unit
      use system, only: system_t
      type(system_t), intent(inout) :: sys
      character(*), intent(out) :: calcname
      real(c_double) :: s(6), dconst
      real(p) :: e_T, e_TT, D_T, D_TT, e_CR, e_CRT
      write(iunit, '(1X, 10("-"))'); write(iunit, '(1X, A)') 'CCSD(T)'; write(iunit, '(1X, 10("-"))')
      call check(afesp_gpu_ccsd_t_spatial(gpu_handle, merge(1_c_int, 0_c_int, sys%ccsd_t_paren), &
                 merge(1_c_int, 0_c_int, sys%ccsd_t_renorm), merge(1_c_int, 0_c_int, sys%ccsd_t_comp_renorm), s, dconst), 'ccsd_t')
      e_T = s(1); e_TT = s(2); D_T = s(3); D_TT = s(4); e_CR = s(5); e_CRT = s(6)
      if (sys%ccsd_t_renorm .or. sys%ccsd_t_comp_renorm) then
         D_T = D_T + dconst
         if (sys%ccsd_t_paren) D_TT = D_TT + dconst
      end if
      sys%e_ccsd_t = sys%e_ccsd + e_T
      sys%e_highest = sys%e_ccsd_t
      if (sys%ccsd_t_paren) then
         sys%e_ccsd_tt = sys%e_ccsd + e_TT
         sys%e_highest = sys%e_ccsd_tt
      end if
      if (sys%ccsd_t_renorm .or. sys%ccsd_t_comp_renorm) then
         sys%e_rccsd_t = sys%e_ccsd + e_T/D_T
         sys%e_highest = sys%e_rccsd_t
         sys%D_T = D_T
         if (sys%ccsd_t_paren) then
            sys%e_rccsd_tt = sys%e_ccsd + e_TT/D_TT
            sys%e_highest = sys%e_rccsd_tt
         end if
         if (sys%ccsd_t_comp_renorm) then
            sys%e_crccsd_t = sys%e_ccsd + e_CR/D_T
            sys%e_highest = sys%e_crccsd_t
            sys%D_TT = D_TT
            if (sys%ccsd_t_paren) then
               sys%e_crccsd_tt = sys%e_ccsd + e_CRT/D_TT
               sys%e_highest = sys%e_crccsd_tt
            end if
         end if
      end if
      calcname = 'CCSD'
      if (sys%ccsd_t_paren) then
         calcname = trim(calcname)//'(T)'
      else
         calcname = trim(calcname)//'[T]'
      end if
      if (sys%ccsd_t_renorm) calcname = 'renormalised '//trim(calcname)
      if (sys%ccsd_t_comp_renorm) calcname = 'completely renormalised '//trim(calcname)
      write(iunit, '(1X, A, 1X, F15.9)') 'Restricted '//trim(calcname)//' correlation energy (Hartree):', sys%e_highest
   end subroutine gpu_ccsd_t_spatial

   subroutine gpu_ccsd_t_spinorb(sys)
      ! Replaces do_ccsd_t_spinorb / do_ccsd_t_spinorb_acc (src/ccsd.f90:1812-2016)
      use, intrinsic :: iso_fortran_env, only: iunit => output_unit
      use system, only: system_t
      type(system_t), intent(inout) :: sys
      real(c_double) :: e_T
      write(iunit, '(1X, 10("-"))'); write(iunit, '(1X, A)') 'CCSD(T)'; write(iunit, '(1X, 10("-"))')
      call check(afesp_gpu_ccsd_t_spinorb(gpu_handle, e_T), 'ccsd_t_spinorb')
      sys%e_ccsd_t = e_T + sys%e_ccsd
      sys%e_highest = sys%e_ccsd_t
      write(iunit, '(1X, A, 1X, F15.9)') 'Unrestricted CCSD(T) correlation energy (Hartree):', sys%e_ccsd_t
   end subroutine gpu_ccsd_t_spinorb

end module afesp_gpu
