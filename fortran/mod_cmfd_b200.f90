!> Drop-in replacement of ADPRES 1.2 src/mod_cmfd.f90: same module name, same public
!! procedures and argument lists (outer, outer_fs, outer_ad, outer_th, outer_tr, PowDis,
!! Integrate, ...), bodies call the CUDA hot path through fortran/adpres_b200_bind.f90.
!! The `do p = 1, nout` loops, the per-iteration prints, the exit tests and the STOP messages
!! stay here, exactly as in the reference (src/mod_cmfd.f90:465-509 etc.); one pass of the loop
!! body is one adp_outer_iter() call (a CUDA-graph replay + a read-back of three scalars).
module CMFD

  use sdata, only: dp
  use adpres_b200
  implicit none
  save

contains

  !> nodal_upd(popt, nmode), src/mod_cmfd.f90:339-383
  subroutine nodal_upd(popt, nmode)
    use sdata, only: ndmax, im, jm, km, get_time, nod_time
    use io,    only: ounit, scr
    integer, intent(in) :: popt, nmode
    real(dp) :: st, fn
    integer(c_int) :: ierr
    st = get_time()
    ierr = adp_nodal_upd(ctx, nmode, ndmax, im, jm, km)   ! ndmax = 0; nodal_update[_pnm]; matrix_setup(0)
    call gpu_check(ierr, 'adp_nodal_upd')
    if (ierr == 2) then
      write(*,*) 'ERROR IN MATRIX DECOMP: DIAGONAL ELEMENTS CLOSE TO ZERO'
      stop
    else if (ierr == 3) then
      write(*,1236) ndmax
      write(*,*) "The two-node nonlinear iteration seems not stable."
      stop
    end if
    if (popt > 0) then
      write(ounit,*) '    .....NODAL COUPLING UPDATED..... '
      write(ounit,1145) ndmax, im, jm, km
      if (scr) then
        write(*,*) '    .....NODAL COUPLING UPDATED..... '
        write(*,1145) ndmax, im, jm, km
      end if
    end if
    fn = get_time()
    nod_time = nod_time + (fn-st)
    1145 format ('MAX. CHANGE IN NODAL COUPLING COEF.= ', ES12.5, ' AT NODE I = ', I2, ', J = ', I2, ', K = ', I2)
    1236 format(" Error: Max. change in nodal coupling coefficient = ", F10.1)
  end subroutine nodal_upd

  subroutine print_keff(popt)
    use sdata, only: Ke
    use io,    only: ounit, scr
    integer, intent(in) :: popt
    if (popt > 0) then
      write(ounit,*); write(ounit,1146) Ke
      if (scr) then; write(*,*); write(*,1146) Ke; end if
    end if
    1146 format(2X,'MULTIPLICATION EFFECTIVE (K-EFF) = ', F9.6)
  end subroutine print_keff

  !> common driver of outer / outer_fs / outer_ad: kind 0/1/2.  `first` is the caller's own SAVEd first-call flag
  !! (the reference keeps one per procedure: src/mod_cmfd.f90:439,537,626).
  subroutine outer_common(kind, popt, label, first)
    use sdata, only: ng, nnod, nout, serc, ferc, fer, ser, f0, fs0, s0, nupd, Ke, nac, ndmax, kern, &
                     get_time, fdm_time
    use io,    only: ounit, scr, bther
    integer, intent(in) :: kind, popt
    character(len=*), intent(in) :: label
    logical, intent(inout) :: first
    integer :: p, mode, nmode
    integer(c_int) :: ierr
    logical :: init
    real(dp) :: st, fn

    st = get_time()
    call gpu_push_inputs(kind == 1)                              ! exsrc (%ESRC) only matters to outer_fs
    ierr = adp_matrix_setup(ctx, 1_c_int)                       ! CALL matrix_setup(1)
    call gpu_check(ierr, 'adp_matrix_setup')
    mode = ADP_MODE_FORWARD; nmode = 1
    if (kind == 1) mode = ADP_MODE_FIXEDSRC
    if (kind == 2) then; mode = ADP_MODE_ADJOINT; nmode = 0; end if
    ! first-call initialisation: outer `first .and. bther == 0` (:448), outer_fs `first` (:544), outer_ad `first .and. popt > 0` (:635)
    init = first
    if (kind == 0) init = first .and. bther == 0
    if (kind == 2) init = first .and. popt > 0
    if (init) then
      allocate (f0(nnod,ng), fs0(nnod), s0(nnod,ng))
      s0 = 0._dp                                                 ! stays on the device (see gpu_pull_results)
      if (kind == 2) then
        ierr = adp_init_flux(ctx, 1_c_int)                      ! Ke = 1; f0 = 1; FSrcAd(fs0)
      else
        ierr = adp_init_flux(ctx, 0_c_int)                      ! Ke = 1; f0 = 1; FSrc(fs0)
      end if
      call gpu_check(ierr, 'adp_init_flux')
      first = .false.
    end if
    ierr = adp_outer_begin(ctx, mode)                            ! f = Integrate(fs0); e1 = Integrate(errn = 1)
    call gpu_check(ierr, 'adp_outer_begin')
    fn = get_time()
    fdm_time = fdm_time + (fn-st)

    do p = 1, nout
      st = get_time()
      ierr = adp_outer_iter(ctx, mode, p, Ke, ser, fer)          ! src/mod_cmfd.f90:467-487 on the GPU
      call gpu_check(ierr, 'adp_outer_iter')
      if (MOD(p,nac) == 0 .and. popt > 0) then
        write(ounit,*) '    ...FISSION SOURCE EXTRAPOLATED...'
        if (scr) write(*,*) '    ...FISSION SOURCE EXTRAPOLATED...'
      end if
      fn = get_time()
      fdm_time = fdm_time + (fn-st)
      if (MOD(p,nupd) == 0 .and. kern /= ' FDM' .and. .not. (kind == 2 .and. popt <= 0)) call nodal_upd(popt, nmode)
      if (popt > 0) then
        if (kind == 1) then
          write(ounit,'(I5,2ES15.5)') p, ser, fer
          if (scr) write(*,'(I5,2ES15.5)') p, ser, fer
        else
          write(ounit,'(I5,F13.6,2ES15.5)') p, Ke, ser, fer
          if (scr) write(*,'(I5,F13.6,2ES15.5)') p, Ke, ser, fer
        end if
      end if
      if ((ser < serc) .AND. (fer < ferc) .AND. (ndmax < 1.e-2)) exit
    end do

    if (p-1 == nout) THEN
      write(*,*)
      write(*,*) '  MAXIMUM NUMBER OF OUTER ITERATION IS REACHED IN '//label//' CALCULATION.'
      write(*,*) '  CHECK PROBLEM SPECIFICATION OR CHANGE ITERATION CONTROL (%ITER).'
      write(*,*) '  PERHAPS BY MAKING FISSION SOURCE INTERPOLATION MORE FREQUENT'
      write(*,*) '  ADPRES IS STOPING...'
      STOP
    end if
    call gpu_pull_results()                                       ! f0, fs0, Ke (and nod for RODEJECT) -> sdata
    if (kind /= 1) call print_keff(popt)
  end subroutine outer_common

  subroutine outer(popt)          ! src/mod_cmfd.f90:415-509
    integer, optional, intent(in) :: popt
    logical, save :: first = .true.
    call outer_common(0, popt, 'FORWARD', first)
  end subroutine outer

  subroutine outer_fs(popt)       ! src/mod_cmfd.f90:513-598
    integer, optional, intent(in) :: popt
    logical, save :: first = .true.
    call outer_common(1, popt, 'FIXED-SOURCE', first)
  end subroutine outer_fs

  subroutine outer_ad(popt)       ! src/mod_cmfd.f90:602-699
    integer, optional, intent(in) :: popt
    logical, save :: first = .true.
    call outer_common(2, popt, 'ADJOINT', first)
  end subroutine outer_ad

  !> outer_th(maxn), src/mod_cmfd.f90:703-796
  subroutine outer_th(maxn)
    use sdata, only: ng, nnod, serc, ferc, fer, ser, f0, fs0, s0, nupd, Ke, ndmax, nth, kern, get_time, fdm_time
    use io,    only: ounit, biter
    integer, intent(in) :: maxn
    integer :: p
    integer(c_int) :: ierr
    logical, save :: first = .true.
    logical :: lnupd
    lnupd = .true.
    call gpu_push_inputs(.false.)
    ierr = adp_matrix_setup(ctx, 1_c_int)
    call gpu_check(ierr, 'adp_matrix_setup')
    if (first) then
      allocate (f0(nnod,ng), fs0(nnod), s0(nnod,ng))
      s0 = 0._dp
      ierr = adp_init_flux(ctx, 0_c_int)
      call gpu_check(ierr, 'adp_init_flux')
      first = .false.
    end if
    ierr = adp_outer_begin(ctx, ADP_MODE_FORWARD)
    call gpu_check(ierr, 'adp_outer_begin')
    if (biter == 0) nupd = int(nth/2)
    do p = 1, maxn
      ierr = adp_outer_iter(ctx, ADP_MODE_FORWARD, p, Ke, ser, fer)
      call gpu_check(ierr, 'adp_outer_iter')
      if (MOD(p,nupd) == 0 .and. kern /= ' FDM') then
        lnupd = .false.
        call nodal_upd(0, 1)
      end if
      if ((ser < serc) .AND. (fer < ferc) .AND. (ndmax < 1.e-2)) exit
    end do
    if (lnupd .and. kern /= ' FDM') then
      write(*,*) 'ERROR: OUTER ITERATION WITHIN T-H ITERATION FINISHED WITHOUT NODAL UPDATE'
      write(*,*) 'CHANGE ITERATION CONTROL USING %ITER CARD'
      stop
    end if
    call gpu_pull_results()
  end subroutine outer_th

  !> outer_tr(ht, maxi), src/mod_cmfd.f90:800-868.  trans_calc has put sigr (+1/(theta v dt)+omega/v),
  !! sigrp, ft, fst into sdata (src/mod_trans.f90:398-416); c0, omeg, L come from the previous step.
  subroutine outer_tr(ht, maxi)
    use sdata, only: nout, serc, ferc, fer, ser, nupd, Ke, ndmax, kern, exsrc, dfis, c0, ft, fst, omeg, sigrp, L, &
                     ibeta, lamb, velo, tbeta, sth, bth, m, nmat, ng, nf
    use io, only: bxtab
    real(dp), intent(in)  :: ht
    logical, intent(out)  :: maxi
    integer :: p, i
    integer(c_int) :: ierr
    real(c_double), allocatable :: mib(:,:), mla(:,:), mve(:,:)
    call gpu_push_inputs(.false.)                                 ! exsrc is computed on the device below
    if (bxtab == 1) then        ! %XTAB decks: kinetics data per material (get_exsrc, src/mod_cmfd.f90:898-925)
      allocate(mib(nf,nmat), mla(nf,nmat), mve(ng,nmat))
      do i = 1, nmat
        mib(:,i) = m(i)%iBeta; mla(:,i) = m(i)%lamb; mve(:,i) = m(i)%velo
      end do
      ierr = adp_set_kinetics_xtab(ctx, mib, mla, mve, tbeta, sth, bth)
    else
      ierr = adp_set_kinetics(ctx, ibeta, lamb, velo, tbeta, sth, bth)
    end if
    call gpu_check(ierr, 'adp_set_kinetics')
    ierr = adp_set_transient(ctx, c0, ft, fst, omeg, sigrp, L)
    call gpu_check(ierr, 'adp_set_transient')
    ierr = adp_matrix_setup(ctx, 1_c_int)
    call gpu_check(ierr, 'adp_matrix_setup')
    ierr = adp_get_exsrc(ctx, ht)                                 ! get_exsrc(ht, exsrc) on the device
    call gpu_check(ierr, 'adp_get_exsrc')
    ierr = adp_outer_begin(ctx, ADP_MODE_TRANSIENT)
    call gpu_check(ierr, 'adp_outer_begin')
    do p = 1, nout
      ierr = adp_outer_iter(ctx, ADP_MODE_TRANSIENT, p, Ke, ser, fer)
      call gpu_check(ierr, 'adp_outer_iter')
      if (MOD(p,nupd) == 0 .and. kern /= ' FDM') call nodal_upd(0, 2)
      if ((ser < serc) .AND. (fer < ferc) .AND. (ndmax < 1.e-2)) exit
    end do
    maxi = (p == nout+1)
    ierr = adp_get_exsrc_arrays(ctx, exsrc, dfis)                  ! uPden / the next step's glue read dfis, exsrc
    call gpu_check(ierr, 'adp_get_exsrc_arrays')
    call gpu_pull_results()
  end subroutine outer_tr

  !> PowDis(p), src/mod_cmfd.f90:1290-1333
  subroutine PowDis(p)
    use sdata, only: mode
    use io,    only: ounit
    real(dp), dimension(:), intent(out) :: p
    integer(c_int) :: ierr, fs
    fs = 0
    if (mode == 'FIXEDSRC') fs = 1
    ierr = adp_powdis(ctx, p, fs)
    call gpu_check(ierr, 'adp_powdis')
    if (ierr == 4) then
      write(ounit, *) '   ERROR: TOTAL NODES POWER IS ZERO OR LESS'
      write(ounit, *) '   STOP IN subroutine POWDIS'
      STOP
    end if
  end subroutine PowDis

  !> Integrate(s), src/mod_cmfd.f90:1120-1139
  function Integrate(s) result(intg)
    real(dp), dimension(:), intent(in) :: s
    real(dp) :: intg
    integer(c_int) :: ierr
    ierr = adp_integrate(ctx, s, intg)
    call gpu_check(ierr, 'adp_integrate')
  end function Integrate

end module CMFD
