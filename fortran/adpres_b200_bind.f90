!> ISO_C_BINDING interface to libadpres_b200.so (include/adpres_b200.h).
!!
!! This module and fortran/mod_cmfd_b200.f90 are what a maintainer adds to ADPRES 1.2 to run
!! the CMFD/BiCGSTAB + SANM/PNM hot path on a B200:  src/mod_cmfd.f90 is replaced by
!! fortran/mod_cmfd_b200.f90 (same module name `cmfd`, same public procedures and argument
!! lists), every other source file -- ADPRES.f90, mod_io.f90 (deck parser), mod_data.f90,
!! mod_xsec.f90, mod_nodal.f90 (Lxyz is still used on the host by `reactivity`), mod_th.f90,
!! mod_trans.f90, mod_control.f90 -- stays byte-for-byte as it is.
!!
!! NOTE: no Fortran compiler exists in the image this was developed in (gfortran, flang, ifx,
!! nvfortran all absent), so these two files are delivered as source and are exercised only
!! indirectly: the C ABI they bind is driven by the identical call sequence from
!! adpres_b200/csrc/host_cmfd.cpp and from the Python ctypes binding in the tests.
module adpres_b200
  use iso_c_binding
  implicit none

  integer(c_int), parameter :: ADP_MODE_FORWARD = 0, ADP_MODE_ADJOINT = 1, &
                               ADP_MODE_FIXEDSRC = 2, ADP_MODE_TRANSIENT = 3
  integer(c_int), parameter :: ADP_KERN_FDM = 0, ADP_KERN_PNM = 1, ADP_KERN_SANM = 2

  type(c_ptr), save :: ctx = c_null_ptr      ! one device context per process (the reference is SAVE'd and serial)

  interface
    integer(c_int) function adp_create(ctx, device) bind(C, name="adp_create")
      import; type(c_ptr), intent(out) :: ctx; integer(c_int), value :: device
    end function
    integer(c_int) function adp_destroy(ctx) bind(C, name="adp_destroy")
      import; type(c_ptr), value :: ctx
    end function
    !> multi-GPU (one process per GPU): reads ADP_NRANKS / ADP_RANK / ADP_UID_FILE from the environment
    integer(c_int) function adp_comm_init_env(ctx) bind(C, name="adp_comm_init_env")
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function adp_set_geometry(ctx, nxx, nyy, nzz, nnod, ng, nmat, ix, iy, iz, ysmin, ysmax, &
                                             xsmin, xsmax, xdel, ydel, zdel, bc, mat) bind(C, name="adp_set_geometry")
      import; type(c_ptr), value :: ctx
      integer(c_int), value :: nxx, nyy, nzz, nnod, ng, nmat
      integer(c_int), intent(in) :: ix(*), iy(*), iz(*), ysmin(*), ysmax(*), xsmin(*), xsmax(*), bc(6), mat(*)
      real(c_double), intent(in) :: xdel(*), ydel(*), zdel(*)
    end function
    integer(c_int) function adp_set_xs(ctx, D, sigr, nuf, sigf, sigs, chi, dc, exsrc) bind(C, name="adp_set_xs")
      import; type(c_ptr), value :: ctx
      real(c_double), intent(in) :: D(*), sigr(*), nuf(*), sigf(*), sigs(*), chi(*), dc(*), exsrc(*)
    end function
    !> adp_set_xs reading only the arrays whose bit is set in mask (0 D, 1 sigr, 2 nuf, 3 sigf, 4 sigs, 5 chi, 6 dc, 7 exsrc)
    integer(c_int) function adp_set_xs_mask(ctx, mask, D, sigr, nuf, sigf, sigs, chi, dc, exsrc) bind(C, name="adp_set_xs_mask")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: mask
      real(c_double), intent(in) :: D(*), sigr(*), nuf(*), sigf(*), sigs(*), chi(*), dc(*), exsrc(*)
    end function
    integer(c_int) function adp_set_control(ctx, nout, nin, nac, nupd, serc, ferc, kern) bind(C, name="adp_set_control")
      import; type(c_ptr), value :: ctx
      integer(c_int), value :: nout, nin, nac, nupd, kern
      real(c_double), value :: serc, ferc
    end function
    integer(c_int) function adp_matrix_setup(ctx, opt) bind(C, name="adp_matrix_setup")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: opt
    end function
    integer(c_int) function adp_init_flux(ctx, adjoint) bind(C, name="adp_init_flux")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: adjoint
    end function
    integer(c_int) function adp_outer_begin(ctx, mode) bind(C, name="adp_outer_begin")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: mode
    end function
    integer(c_int) function adp_outer_iter(ctx, mode, p, Ke, ser, fer) bind(C, name="adp_outer_iter")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: mode, p
      real(c_double), intent(out) :: Ke, ser, fer
    end function
    integer(c_int) function adp_nodal_upd(ctx, nmode, ndmax, im, jm, km) bind(C, name="adp_nodal_upd")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nmode
      real(c_double), intent(out) :: ndmax; integer(c_int), intent(out) :: im, jm, km
    end function
    integer(c_int) function adp_powdis(ctx, p, fixedsrc_mode) bind(C, name="adp_powdis")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: p(*); integer(c_int), value :: fixedsrc_mode
    end function
    integer(c_int) function adp_integrate(ctx, s, res) bind(C, name="adp_integrate")
      import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: s(*); real(c_double), intent(out) :: res
    end function
    integer(c_int) function adp_set_kinetics_xtab(ctx, mibeta, mlamb, mvelo, tbeta, sth, bth) bind(C, name="adp_set_kinetics_xtab")
      import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: mibeta(6,*), mlamb(6,*), mvelo(*), tbeta(*)
      real(c_double), value :: sth, bth
    end function
    integer(c_int) function adp_set_kinetics(ctx, ibeta, lamb, velo, tbeta, sth, bth) bind(C, name="adp_set_kinetics")
      import; type(c_ptr), value :: ctx
      real(c_double), intent(in) :: ibeta(6), lamb(6), velo(*), tbeta(*); real(c_double), value :: sth, bth
    end function
    integer(c_int) function adp_set_transient(ctx, c0, ft, fst, omeg, sigrp, L) bind(C, name="adp_set_transient")
      import; type(c_ptr), value :: ctx
      real(c_double), intent(in) :: c0(*), ft(*), fst(*), omeg(*), sigrp(*), L(*)
    end function
    integer(c_int) function adp_get_exsrc(ctx, ht) bind(C, name="adp_get_exsrc")
      import; type(c_ptr), value :: ctx; real(c_double), value :: ht
    end function
    integer(c_int) function adp_get_exsrc_arrays(ctx, exsrc, dfis) bind(C, name="adp_get_exsrc_arrays")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: exsrc(*), dfis(*)
    end function
    integer(c_int) function adp_get_state(ctx, f0, fs0, s0, Ke) bind(C, name="adp_get_state")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: f0(*), fs0(*), s0(*), Ke
    end function
    !> adp_get_state writing only the arrays whose bit is set in mask (0 f0, 1 fs0, 2 s0)
    integer(c_int) function adp_get_state_mask(ctx, mask, f0, fs0, s0, Ke) bind(C, name="adp_get_state_mask")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: mask
      real(c_double), intent(out) :: f0(*), fs0(*), s0(*), Ke
    end function
    type(c_ptr) function adp_last_error(ctx) bind(C, name="adp_last_error")
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function adp_set_state(ctx, f0, fs0, Ke) bind(C, name="adp_set_state")
      import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: f0(*), fs0(*); real(c_double), value :: Ke
    end function
    integer(c_int) function adp_set_s0(ctx, s0, g) bind(C, name="adp_set_s0")
      import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: s0(*); integer(c_int), value :: g
    end function
    integer(c_int) function adp_get_nod(ctx, df, dn) bind(C, name="adp_get_nod")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: df(*), dn(*)
    end function
    integer(c_int) function adp_set_nod_dn(ctx, dn) bind(C, name="adp_set_nod_dn")
      import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: dn(*)
    end function
    integer(c_int) function adp_get_ndmax(ctx, ndmax) bind(C, name="adp_get_ndmax")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: ndmax
    end function
    ! ---- optional: time-step glue of mod_trans.f90 and XS_updt on the device (INTEGRATION.md) ----
    integer(c_int) function adp_lxyz_total(ctx, L) bind(C, name="adp_lxyz_total")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: L(*)
    end function
    integer(c_int) function adp_save_adjoint(ctx) bind(C, name="adp_save_adjoint")
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function adp_ipden(ctx) bind(C, name="adp_ipden")
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function adp_update_omeg(ctx, ht, bextr) bind(C, name="adp_update_omeg")
      import; type(c_ptr), value :: ctx; real(c_double), value :: ht; integer(c_int), value :: bextr
    end function
    integer(c_int) function adp_begin_time_step(ctx, ht) bind(C, name="adp_begin_time_step")
      import; type(c_ptr), value :: ctx; real(c_double), value :: ht
    end function
    integer(c_int) function adp_upden(ctx, ht) bind(C, name="adp_upden")
      import; type(c_ptr), value :: ctx; real(c_double), value :: ht
    end function
    integer(c_int) function adp_powtot(ctx, tpow) bind(C, name="adp_powtot")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: tpow
    end function
    integer(c_int) function adp_reactivity(ctx, use_sigrp, rho) bind(C, name="adp_reactivity")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: use_sigrp; real(c_double), intent(out) :: rho
    end function
    integer(c_int) function adp_xs_update(ctx, bpos) bind(C, name="adp_xs_update")
      import; type(c_ptr), value :: ctx; real(c_double), intent(in) :: bpos(*)
    end function
    integer(c_int) function adp_set_feedback(ctx, which, ref, dsigtr, dsiga, dnuf, dsigf, dsigs) bind(C, name="adp_set_feedback")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: which; real(c_double), value :: ref
      real(c_double), intent(in) :: dsigtr(*), dsiga(*), dnuf(*), dsigf(*), dsigs(*)
    end function
    integer(c_int) function adp_xs_update_th(ctx, bcon, ftem, mtem, cden, bpos) bind(C, name="adp_xs_update_th")
      import; type(c_ptr), value :: ctx, ftem, mtem, cden      ! c_null_ptr = the TH state on the device
      real(c_double), value :: bcon; real(c_double), intent(in) :: bpos(*)
    end function
    ! %XTAB decks: the branch tables of sdata's m(1:nmat) (MBRANCH), packed by the caller -- see include/adpres_b200.h
    integer(c_int) function adp_set_xtab(ctx, dims, trod, par, xs, rxs) bind(C, name="adp_set_xtab")
      import; type(c_ptr), value :: ctx, rxs                    ! rxs: c_null_ptr if no material has a rodded set
      integer(c_int), intent(in) :: dims(4,*), trod(*); real(c_double), intent(in) :: par(*), xs(*)
    end function
    integer(c_int) function adp_set_crod_map(ctx, nb, pos0, ssize, fbmap) bind(C, name="adp_set_crod_map")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nb; real(c_double), value :: pos0, ssize
      integer(c_int), intent(in) :: fbmap(*)
    end function
    integer(c_int) function adp_xs_update_xtab(ctx, bcon, ftem, mtem, cden, bpos) bind(C, name="adp_xs_update_xtab")
      import; type(c_ptr), value :: ctx, ftem, mtem, cden      ! c_null_ptr = the TH state on the device
      real(c_double), value :: bcon; real(c_double), intent(in) :: bpos(*)
    end function
    integer(c_int) function adp_get_dc(ctx, dc) bind(C, name="adp_get_dc")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: dc(*)
    end function
    integer(c_int) function adp_get_errors(ctx, ser, fer) bind(C, name="adp_get_errors")
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: ser, fer
    end function
    integer(c_int) function adp_set_th(ctx, pi, rf, rg, rc, dia, dh, farea, cflow, cf, tin, rpos, rdel, ntem, stab) &
        bind(C, name="adp_set_th")
      import; type(c_ptr), value :: ctx
      real(c_double), value :: pi, rf, rg, rc, dia, dh, farea, cflow, cf, tin
      real(c_double), intent(in) :: rpos(*), rdel(*), stab(*); integer(c_int), value :: ntem
    end function
    integer(c_int) function adp_set_th_state(ctx, tfm, heatf, ent, ftem, mtem, cden, frate) bind(C, name="adp_set_th_state")
      import; type(c_ptr), value :: ctx, frate          ! frate: c_null_ptr before the first transient step
      real(c_double), intent(in) :: tfm(*), heatf(*), ent(*), ftem(*), mtem(*), cden(*)
    end function
    integer(c_int) function adp_get_th_state(ctx, tfm, heatf, ent, ftem, mtem, cden, frate) bind(C, name="adp_get_th_state")
      import; type(c_ptr), value :: ctx
      real(c_double), intent(out) :: tfm(*), heatf(*), ent(*), ftem(*), mtem(*), cden(*), frate(*)
    end function
    integer(c_int) function adp_th_pline(ctx, pow, ppow, form, node_nf) bind(C, name="adp_th_pline")
      import; type(c_ptr), value :: ctx; real(c_double), value :: pow, ppow; integer(c_int), value :: form
      real(c_double), intent(in) :: node_nf(*)
    end function
    integer(c_int) function adp_th_upd(ctx, xpline, th_err) bind(C, name="adp_th_upd")
      import; type(c_ptr), value :: ctx, xpline         ! xpline: c_null_ptr = the device's pline
      real(c_double), intent(out) :: th_err
    end function
    integer(c_int) function adp_th_trans(ctx, xpline, h) bind(C, name="adp_th_trans")
      import; type(c_ptr), value :: ctx, xpline; real(c_double), value :: h
    end function
  end interface

contains

  !> First call: create the context and hand over the geometry (what inp_geom1/2 + misc left in sdata).
  subroutine gpu_init()
    use sdata, only: nxx, nyy, nzz, nnod, ng, nmat, ix, iy, iz, ystag, xstag, xdel, ydel, zdel, &
                     xeast, xwest, ynorth, ysouth, zbott, ztop, mat
    integer(c_int) :: ierr, bc(6), i
    integer(c_int), allocatable :: ysmin(:), ysmax(:), xsmin(:), xsmax(:)
    if (c_associated(ctx)) return
    ierr = adp_create(ctx, -1_c_int)      ! device = LOCAL_RANK (one process per GPU), else 0
    if (ierr /= 0) stop 'adpres_b200: no CUDA device (there is no CPU fallback)'
    ierr = adp_comm_init_env(ctx)         ! ADP_NRANKS|WORLD_SIZE > 1: needs ADP_UID_FILE or ADP_JOB_ID|MASTER_PORT
    call gpu_check(ierr, 'adp_comm_init_env')
    allocate(ysmin(nyy), ysmax(nyy), xsmin(nxx), xsmax(nxx))
    do i = 1, nyy; ysmin(i) = ystag(i)%smin; ysmax(i) = ystag(i)%smax; end do
    do i = 1, nxx; xsmin(i) = xstag(i)%smin; xsmax(i) = xstag(i)%smax; end do
    bc = (/ xeast, xwest, ynorth, ysouth, zbott, ztop /)
    ierr = adp_set_geometry(ctx, nxx, nyy, nzz, nnod, ng, nmat, ix, iy, iz, ysmin, ysmax, xsmin, xsmax, &
                            xdel, ydel, zdel, bc, mat)
    call gpu_check(ierr, 'adp_set_geometry')
  end subroutine gpu_init

  !> Every adp_* call returns 0, a positive code for one of the reference's own STOP conditions (handled where it can
  !! occur) or a negative code for a CUDA / NCCL / usage error.  A negative code is never survivable: the device state is
  !! undefined, so print the library's message and STOP -- never iterate on stale scalars.
  subroutine gpu_check(ierr, what)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    type(c_ptr) :: cmsg
    integer :: i
    if (ierr >= 0) return
    write(*,'(A,A,A,I0)') ' adpres_b200: ', what, ' failed, code ', ierr
    cmsg = adp_last_error(ctx)
    if (c_associated(cmsg)) then
      call c_f_pointer(cmsg, msg, (/ 1024 /))
      do i = 1, 1024
        if (msg(i) == c_null_char) exit
        write(*,'(A)',advance='no') msg(i)
      end do
      write(*,*)
    end if
    stop 'adpres_b200: device error'
  end subroutine gpu_check

  !> Before every outer*(): cross sections as XS_updt left them + the iteration control.
  !! Only what the caller can have changed since the previous outer*() call goes over PCIe:
  !!   first call            everything
  !!   later calls           D, sigr, nuf, sigf, sigs (XS_updt / XStab_updt rewrite them: mod_xsec.f90:11-86)
  !!   dc                    only with %XTAB (XStab_updt takes the ADFs from the tables, mod_xsec.f90:73-75)
  !!   exsrc                 only in outer_fs (%ESRC); outer_tr computes it on the device (get_exsrc, mod_cmfd.f90:830),
  !!                         everywhere else it stays zero
  !!   chi                   never changes after the first call
  subroutine gpu_push_inputs(with_exsrc)
    use sdata, only: D, sigr, nuf, sigf, sigs, chi, dc, exsrc, nout, nin, nac, nupd, serc, ferc, kern
    use io,    only: bxtab
    logical, intent(in) :: with_exsrc
    logical, save :: first = .true.
    integer(c_int) :: ierr, k, mask
    call gpu_init()
    if (first) then
      mask = 255_c_int
      first = .false.
    else
      mask = 31_c_int                              ! D, sigr, nuf, sigf, sigs
      if (bxtab == 1) mask = mask + 64_c_int       ! dc
      if (with_exsrc) mask = mask + 128_c_int      ! exsrc
    end if
    ierr = adp_set_xs_mask(ctx, mask, D, sigr, nuf, sigf, sigs, chi, dc, exsrc)
    call gpu_check(ierr, 'adp_set_xs_mask')
    k = ADP_KERN_SANM
    if (kern == ' FDM') k = ADP_KERN_FDM
    if (kern == ' PNM') k = ADP_KERN_PNM
    ierr = adp_set_control(ctx, nout, nin, nac, nupd, serc, ferc, k)
    call gpu_check(ierr, 'adp_set_control')
  end subroutine gpu_push_inputs

  !> After an outer*(): results the drivers read from sdata -- f0, fs0, Ke always; nod%df/dn only for the
  !! rod-ejection drivers, whose `reactivity` calls Lxyz of the (unchanged) nodal module on the host
  !! (mod_trans.f90:677).  s0 has no reader outside get_exsrc, which runs on the device, and stays there.
  !! On several ranks the library completes every array with the other ranks' slabs (include/adpres_b200.h), so the
  !! host code that follows sees whole arrays on every rank.
  subroutine gpu_pull_results()
    use sdata, only: f0, fs0, s0, Ke, nod, nnod, ng, mode
    integer(c_int) :: ierr
    real(c_double), allocatable :: df(:,:,:), dn(:,:,:)
    integer :: n, g
    ierr = adp_get_state_mask(ctx, 3_c_int, f0, fs0, s0, Ke)
    call gpu_check(ierr, 'adp_get_state_mask')
    if (mode == 'RODEJECT') then
      allocate(df(6,nnod,ng), dn(6,nnod,ng))
      ierr = adp_get_nod(ctx, df, dn)
      call gpu_check(ierr, 'adp_get_nod')
      do g = 1, ng
        do n = 1, nnod
          nod(n,g)%df = df(:,n,g)
          nod(n,g)%dn = dn(:,n,g)
        end do
      end do
      deallocate(df, dn)
    end if
  end subroutine gpu_pull_results

  !> Optional, %XTAB decks: hand the branch tables m(1:nmat) (MBRANCH / XBRANCH, mod_data.f90:176-192, read by
  !! inp_xtab) to the device once; afterwards th_iter / trans_calc replace
  !!     CALL XStab_updt(bcon, ftem, mtem, cden, bpos)                      (mod_th.f90:45-46, mod_trans.f90:393-394)
  !! by ierr = adp_xs_update_xtab(ctx, bcon, c_null_ptr, c_null_ptr, c_null_ptr, bpos)   ! TH fields already on the device
  !! Packing: per branch point [sigtr(ng), siga(ng), nuf(ng), sigf(ng), sigs(g,h) g slow, dc(g,face) g slow],
  !! points ordered (s,t,u,v) with v (moderator temperature) fastest -- see include/adpres_b200.h.
  subroutine gpu_set_xtab()
    use sdata, only: ng, nmat, m, nb, pos0, ssize, fbmap
    use io, only: bcrod
    integer(c_int) :: ierr
    integer(c_int), allocatable :: dims(:,:), trod(:)
    real(c_double), allocatable, target :: par(:), xs(:), rxs(:)
    integer :: i, s, t, u, v, g, h, k, nval, npar, ntab, ip, it
    logical :: anyrod
    call gpu_init()
    nval = 4*ng + ng*ng + 6*ng
    allocate(dims(4,nmat), trod(nmat))
    npar = 0; ntab = 0; anyrod = .false.
    do i = 1, nmat
      dims(:,i) = (/ m(i)%nd, m(i)%nb, m(i)%nf, m(i)%nm /)
      trod(i) = m(i)%trod
      npar = npar + m(i)%nd + m(i)%nb + m(i)%nf + m(i)%nm
      ntab = ntab + m(i)%nd * m(i)%nb * m(i)%nf * m(i)%nm * nval
      if (m(i)%trod == 1) anyrod = .true.
    end do
    allocate(par(npar), xs(ntab), rxs(ntab))
    rxs = 0._c_double
    ip = 0; it = 0
    do i = 1, nmat
      par(ip+1:ip+m(i)%nd) = m(i)%pd(1:m(i)%nd); ip = ip + m(i)%nd
      par(ip+1:ip+m(i)%nb) = m(i)%pb(1:m(i)%nb); ip = ip + m(i)%nb
      par(ip+1:ip+m(i)%nf) = m(i)%pf(1:m(i)%nf); ip = ip + m(i)%nf
      par(ip+1:ip+m(i)%nm) = m(i)%pm(1:m(i)%nm); ip = ip + m(i)%nm
      do s = 1, m(i)%nd; do t = 1, m(i)%nb; do u = 1, m(i)%nf; do v = 1, m(i)%nm
        call pack_branch(m(i)%xsec(s,t,u,v), xs(it+1:it+nval))
        if (m(i)%trod == 1) call pack_branch(m(i)%rxsec(s,t,u,v), rxs(it+1:it+nval))
        it = it + nval
      end do; end do; end do; end do
    end do
    if (anyrod) then
      ierr = adp_set_xtab(ctx, dims, trod, par, xs, c_loc(rxs))
    else
      ierr = adp_set_xtab(ctx, dims, trod, par, xs, c_null_ptr)
    end if
    call gpu_check(ierr, 'adp_set_xtab')
    if (bcrod == 1) then
      ierr = adp_set_crod_map(ctx, nb, pos0, ssize, fbmap)       ! fbmap(nxx,nyy), column-major as the C side expects
      call gpu_check(ierr, 'adp_set_crod_map')
    end if
  contains
    subroutine pack_branch(x, a)
      use sdata, only: XBRANCH
      type(XBRANCH), intent(in) :: x
      real(c_double), intent(out) :: a(:)
      a(1:ng) = x%sigtr; a(ng+1:2*ng) = x%siga; a(2*ng+1:3*ng) = x%nuf; a(3*ng+1:4*ng) = x%sigf
      k = 4*ng
      do g = 1, ng; do h = 1, ng; k = k + 1; a(k) = x%sigs(g,h); end do; end do
      do g = 1, ng; do h = 1, 6;  k = k + 1; a(k) = x%dc(g,h);   end do; end do
    end subroutine pack_branch
  end subroutine gpu_set_xtab

end module adpres_b200
