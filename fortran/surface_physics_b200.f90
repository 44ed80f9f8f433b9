!> ####################################################################
!! **Module**  : surface_physics  (B200 drop-in)
!! **Purpose** : Same module, type and procedure names as mkrapp/semic's
!!               surface_physics.f90, forwarding over ISO_C_BINDING to the
!!               C-ABI of libsemic_b200.so (include/semic_b200.h).
!!
!! Callers (example/example.f90, optimize/run_particles.f90) compile
!! against this file unchanged:  use surface_physics
!!
!! State arrays are POINTER, CONTIGUOUS components associated with the
!! pinned host mirror the library owns (semic_state_field), so user code
!! like `now%sf = forc%sf(:,day)` and `now%tsurf(:) = 260.` keeps
!! working; surface_energy_and_mass_balance uploads the mirror, runs the
!! fused CUDA kernel and downloads the 23 non-forcing fields.
!!
!! NOTE: there is no Fortran compiler in the build image; this file is
!! written against the C header but has not been compiled there.  Build
!! where gfortran exists:  gfortran -c surface_physics_b200.f90
!!                         gfortran example.f90 utils.o surface_physics_b200.o -L<repo>/semic_b200/lib -lsemic_b200
!! ####################################################################
module surface_physics

    use, intrinsic :: iso_c_binding
    implicit none

    integer, parameter:: dp=kind(0.d0)

    ! public constants of the reference module (surface_physics.f90:14-25)
    double precision, parameter :: pi   = 3.141592653589793238462643_dp
    double precision, parameter :: t0   = 273.15_dp
    double precision, parameter :: sigm = 5.67e-8_dp
    double precision, parameter :: eps  = 0.62197_dp
    double precision, parameter :: cls  = 2.83e6_dp
    double precision, parameter :: clm  = 3.30e5_dp
    double precision, parameter :: clv  = 2.5e6_dp
    double precision, parameter :: cap  = 1000.0_dp
    double precision, parameter :: rhow = 1000.0_dp
    double precision, parameter :: hsmax= 5.0_dp
    double precision, parameter :: epsil= epsilon(1.0_dp)

    ! field ids of include/semic_b200.h (declaration order of surface_state_class)
    integer(c_int), parameter :: F_T2M=0, F_TSURF=1, F_HSNOW=2, F_HICE=3, F_ALB=4, F_ALB_SNOW=5, F_MELT=6, &
        F_MELTED_SNOW=7, F_MELTED_ICE=8, F_REFR=9, F_SMB=10, F_ACC=11, F_LHF=12, F_SHF=13, F_LWU=14, F_SUBL=15, &
        F_EVAP=16, F_SMB_SNOW=17, F_SMB_ICE=18, F_RUNOFF=19, F_QMR=20, F_QMR_RES=21, F_AMP=22, F_SF=23, F_RF=24, &
        F_SP=25, F_LWD=26, F_SWD=27, F_WIND=28, F_RHOA=29, F_QQ=30

    !> surface_param_class == struct semic_par (same member order; strings are blank padded)
    type, bind(C) :: surface_param_c
        character(kind=c_char) :: name(256)
        character(kind=c_char) :: boundary(256,30)
        character(kind=c_char) :: alb_scheme(256)
        integer(c_int) :: nx, n_ksub
        real(c_double) :: ceff, albi, albl, alb_smax, alb_smin, hcrit, rcrit, amp, csh, clh, tmin, tmax, &
                          tstic, tsticsub, tau_a, tau_f, w_crit, mcrit, afac, tmid
    end type

    type surface_param_class
        character (len=256) :: name
        character (len=256) :: boundary(30)
        character (len=256) :: alb_scheme
        integer             :: nx
        integer             :: n_ksub
        double precision    :: ceff, albi, albl, alb_smax, alb_smin, hcrit, rcrit, amp, csh, clh, tmin, tmax, &
                               tstic, tsticsub, tau_a, tau_f, w_crit, mcrit, afac, tmid
    end type

    type surface_state_class
        double precision, pointer, contiguous, dimension(:) :: t2m => null(), tsurf => null(), hsnow => null(), &
            hice => null(), alb => null(), alb_snow => null(), melt => null(), melted_snow => null(), &
            melted_ice => null(), refr => null(), smb => null(), acc => null(), lhf => null(), shf => null(), &
            lwu => null(), subl => null(), evap => null(), smb_snow => null(), smb_ice => null(), runoff => null(), &
            qmr => null(), qmr_res => null(), amp => null(), sf => null(), rf => null(), sp => null(), lwd => null(), &
            swd => null(), wind => null(), rhoa => null(), qq => null()
        integer(c_int), pointer, contiguous, dimension(:) :: mask => null()
        type(c_ptr) :: handle = c_null_ptr          !< semic_state* (device SoA + pinned host mirror)
    end type

    !> boundary_opt_class == struct semic_bnd (logical -> c_int 0/1)
    type, bind(C) :: boundary_opt_c
        integer(c_int) :: t2m, tsurf, hsnow, alb, melt, refr, smb, acc, lhf, shf, subl, amp
    end type

    type boundary_opt_class
        logical :: t2m, tsurf, hsnow, alb, melt, refr, smb, acc, lhf, shf, subl, amp
    end type

    type surface_physics_class
        type(surface_param_class) :: par
        type(boundary_opt_class)  :: bnd
        type(boundary_opt_class)  :: bnd0
        type(surface_state_class) :: now
        type(surface_state_class) :: mon
        type(surface_state_class) :: ann
    end type

    interface
        integer(c_int) function semic_surface_alloc(now, npts) bind(C, name="semic_surface_alloc")
            import :: c_ptr, c_int
            type(c_ptr), intent(out) :: now
            integer(c_int), value :: npts
        end function
        integer(c_int) function semic_surface_dealloc(now) bind(C, name="semic_surface_dealloc")
            import :: c_ptr, c_int
            type(c_ptr), value :: now
        end function
        integer(c_int) function semic_surface_physics_par_load(par, filename) bind(C, name="semic_surface_physics_par_load")
            import :: c_int, c_char, surface_param_c
            type(surface_param_c), intent(inout) :: par
            character(kind=c_char), intent(in) :: filename(*)
        end function
        integer(c_int) function semic_surface_boundary_define(bnd, boundary, n, len) bind(C, name="semic_surface_boundary_define")
            import :: c_int, c_char, boundary_opt_c
            type(boundary_opt_c), intent(out) :: bnd
            character(kind=c_char), intent(in) :: boundary(*)
            integer(c_int), value :: n, len
        end function
        integer(c_int) function semic_step(now, par, bnd, day, year) bind(C, name="semic_surface_energy_and_mass_balance")
            import :: c_ptr, c_int, surface_param_c, boundary_opt_c
            type(c_ptr), value :: now
            type(surface_param_c), intent(in) :: par
            type(boundary_opt_c), intent(in) :: bnd
            integer(c_int), value :: day, year
        end function
        integer(c_int) function semic_energy_balance_c(now, par, bnd, day, year) bind(C, name="semic_energy_balance")
            import :: c_ptr, c_int, surface_param_c, boundary_opt_c
            type(c_ptr), value :: now
            type(surface_param_c), intent(in) :: par
            type(boundary_opt_c), intent(in) :: bnd
            integer(c_int), value :: day, year
        end function
        integer(c_int) function semic_mass_balance_c(now, par, bnd, day, year) bind(C, name="semic_mass_balance")
            import :: c_ptr, c_int, surface_param_c, boundary_opt_c
            type(c_ptr), value :: now
            type(surface_param_c), intent(in) :: par
            type(boundary_opt_c), intent(in) :: bnd
            integer(c_int), value :: day, year
        end function
        integer(c_int) function semic_average_c(ave, now, step, nt) bind(C, name="semic_surface_physics_average")
            import :: c_ptr, c_int, c_char, c_double
            type(c_ptr), value :: ave, now
            character(kind=c_char), intent(in) :: step(*)
            real(c_double), value :: nt
        end function
        integer(c_int) function semic_print_param_c(par) bind(C, name="semic_print_param")
            import :: c_int, surface_param_c
            type(surface_param_c), intent(in) :: par
        end function
        integer(c_int) function semic_print_boundary_opt_c(bnd) bind(C, name="semic_print_boundary_opt")
            import :: c_int, boundary_opt_c
            type(boundary_opt_c), intent(in) :: bnd
        end function
        type(c_ptr) function semic_state_field(now, field) bind(C, name="semic_state_field")
            import :: c_ptr, c_int
            type(c_ptr), value :: now
            integer(c_int), value :: field
        end function
        type(c_ptr) function semic_state_mask(now) bind(C, name="semic_state_mask")
            import :: c_ptr
            type(c_ptr), value :: now
        end function
        integer(c_int) function semic_state_download(now, fields) bind(C, name="semic_state_download")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: now
            integer(c_int64_t), value :: fields
        end function
        type(c_ptr) function semic_last_error() bind(C, name="semic_last_error")
            import :: c_ptr
        end function
        ! the three public elementals, evaluated by the library on the device (include/semic_b200.h): over n values for the
        ! rank-1 specifics, one value returned as the function RESULT for the elemental specifics.  Only the latter are
        ! declared PURE: every dummy is passed by VALUE and nothing reaches Fortran through a pointer, so they are pure in
        ! the sense of the standard and an optimising compiler has nothing to assume wrongly.
        integer(c_int) function semic_shf_c(ts, ta, wind, rhoa, csh, cap, shf, n) bind(C, name="semic_sensible_heat_flux")
            import :: c_int, c_double, c_ptr
            real(c_double), intent(in) :: ts(*), ta(*), wind(*), rhoa(*)
            real(c_double), value :: csh, cap
            type(c_ptr), value :: shf
            integer(c_int), value :: n
        end function
        integer(c_int) function semic_lhf_c(ts, wind, shum, sp, rhoatm, mask, clh, eps, cls, clv, lhf, subl, evap, n) &
                bind(C, name="semic_latent_heat_flux")
            import :: c_int, c_double, c_ptr
            real(c_double), intent(in) :: ts(*), wind(*), shum(*), sp(*), rhoatm(*)
            integer(c_int), intent(in) :: mask(*)
            real(c_double), value :: clh, eps, cls, clv
            type(c_ptr), value :: lhf, subl, evap
            integer(c_int), value :: n
        end function
        integer(c_int) function semic_lwu_c(ts, sigma, lwout, n) bind(C, name="semic_longwave_upward")
            import :: c_int, c_double, c_ptr
            real(c_double), intent(in) :: ts(*)
            real(c_double), value :: sigma
            type(c_ptr), value :: lwout
            integer(c_int), value :: n
        end function
        pure real(c_double) function semic_shf_1(ts, ta, wind, rhoa, csh, cap) bind(C, name="semic_sensible_heat_flux_1")
            import :: c_double
            real(c_double), value :: ts, ta, wind, rhoa, csh, cap
        end function
        pure real(c_double) function semic_lhf_1(ts, wind, shum, sp, rhoatm, mask, clh, eps, cls, clv, which) &
                bind(C, name="semic_latent_heat_flux_1")
            import :: c_int, c_double
            real(c_double), value :: ts, wind, shum, sp, rhoatm, clh, eps, cls, clv
            integer(c_int), value :: mask, which
        end function
        pure real(c_double) function semic_lwu_1(ts, sigma) bind(C, name="semic_longwave_upward_1")
            import :: c_double
            real(c_double), value :: ts, sigma
        end function
    end interface

    ! Generic names = the reference's elemental subroutines (surface_physics.f90:378-438).  Both specifics forward to the
    ! C-ABI: the rank-1 specific in ONE call over the whole array (it takes precedence over the elemental one when the
    ! actual arguments are rank-1 arrays, as in energy_balance :173, :183, :191), the elemental one value by value.
    interface sensible_heat_flux
        module procedure sensible_heat_flux_v, sensible_heat_flux_e
    end interface
    interface latent_heat_flux
        module procedure latent_heat_flux_v, latent_heat_flux_e
    end interface
    interface longwave_upward
        module procedure longwave_upward_v, longwave_upward_e
    end interface

    private
    public :: mass_balance, energy_balance, surface_energy_and_mass_balance
    public :: surface_physics_class, surface_state_class, surface_param_class
    public :: boundary_opt_class
    public :: surface_physics_par_load, surface_alloc, surface_dealloc
    public :: surface_boundary_define, surface_physics_average
    public :: print_param, print_boundary_opt
    public :: sigm, cap, eps, cls, clv
    public :: longwave_upward, latent_heat_flux, sensible_heat_flux

contains

    subroutine check(rc, who)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: who
        character(kind=c_char), pointer :: msg(:)
        integer :: n
        if (rc /= 0) then
            call c_f_pointer(semic_last_error(), msg, [1024])
            n = 1
            do while (n < 1024 .and. msg(n) /= c_null_char)
                n = n + 1
            end do
            write(*,*) trim(who), ': ', msg(1:n-1)
            error stop 1
        end if
    end subroutine

    subroutine par_to_c(par, c)
        type(surface_param_class), intent(in) :: par
        type(surface_param_c), intent(out) :: c
        integer :: q, k
        do k = 1, 256
            c%name(k) = par%name(k:k)
            c%alb_scheme(k) = par%alb_scheme(k:k)
            do q = 1, 30
                c%boundary(k,q) = par%boundary(q)(k:k)
            end do
        end do
        c%nx = par%nx; c%n_ksub = par%n_ksub
        c%ceff = par%ceff; c%albi = par%albi; c%albl = par%albl; c%alb_smax = par%alb_smax; c%alb_smin = par%alb_smin
        c%hcrit = par%hcrit; c%rcrit = par%rcrit; c%amp = par%amp; c%csh = par%csh; c%clh = par%clh
        c%tmin = par%tmin; c%tmax = par%tmax; c%tstic = par%tstic; c%tsticsub = par%tsticsub
        c%tau_a = par%tau_a; c%tau_f = par%tau_f; c%w_crit = par%w_crit; c%mcrit = par%mcrit
        c%afac = par%afac; c%tmid = par%tmid
    end subroutine

    subroutine par_from_c(c, par)
        type(surface_param_c), intent(in) :: c
        type(surface_param_class), intent(inout) :: par
        integer :: q, k
        do k = 1, 256
            par%alb_scheme(k:k) = c%alb_scheme(k)
            if (c%alb_scheme(k) == c_null_char) par%alb_scheme(k:k) = ' '
            do q = 1, 30
                par%boundary(q)(k:k) = c%boundary(k,q)
                if (c%boundary(k,q) == c_null_char) par%boundary(q)(k:k) = ' '
            end do
        end do
        par%n_ksub = c%n_ksub
        par%ceff = c%ceff; par%albi = c%albi; par%albl = c%albl; par%alb_smax = c%alb_smax; par%alb_smin = c%alb_smin
        par%hcrit = c%hcrit; par%rcrit = c%rcrit; par%amp = c%amp; par%csh = c%csh; par%clh = c%clh
        par%tmin = c%tmin; par%tmax = c%tmax; par%tstic = c%tstic; par%tsticsub = c%tsticsub
        par%tau_a = c%tau_a; par%tau_f = c%tau_f; par%w_crit = c%w_crit; par%mcrit = c%mcrit
        par%afac = c%afac; par%tmid = c%tmid
    end subroutine

    subroutine bnd_to_c(bnd, c)
        type(boundary_opt_class), intent(in) :: bnd
        type(boundary_opt_c), intent(out) :: c
        c%t2m = merge(1, 0, bnd%t2m); c%tsurf = merge(1, 0, bnd%tsurf); c%hsnow = merge(1, 0, bnd%hsnow)
        c%alb = merge(1, 0, bnd%alb); c%melt = merge(1, 0, bnd%melt); c%refr = merge(1, 0, bnd%refr)
        c%smb = merge(1, 0, bnd%smb); c%acc = merge(1, 0, bnd%acc); c%lhf = merge(1, 0, bnd%lhf)
        c%shf = merge(1, 0, bnd%shf); c%subl = merge(1, 0, bnd%subl); c%amp = merge(1, 0, bnd%amp)
    end subroutine

    !> surface_physics.f90:138-153
    subroutine surface_energy_and_mass_balance(now,par,bnd,day,year)
        type(surface_state_class), intent(in out) :: now
        type(boundary_opt_class),  intent(in)     :: bnd
        type(surface_param_class), intent(in)     :: par
        integer, intent(in) :: day, year
        type(surface_param_c) :: pc
        type(boundary_opt_c)  :: bc
        call par_to_c(par, pc); call bnd_to_c(bnd, bc)
        call check(semic_step(now%handle, pc, bc, int(day, c_int), int(year, c_int)), 'surface_energy_and_mass_balance')
    end subroutine

    !> surface_physics.f90:158-214
    subroutine energy_balance(now,par,bnd,day,year)
        type(surface_state_class), intent(in out) :: now
        type(boundary_opt_class),  intent(in)     :: bnd
        type(surface_param_class), intent(in)     :: par
        integer, intent(in) :: day, year
        type(surface_param_c) :: pc
        type(boundary_opt_c)  :: bc
        call par_to_c(par, pc); call bnd_to_c(bnd, bc)
        call check(semic_energy_balance_c(now%handle, pc, bc, int(day, c_int), int(year, c_int)), 'energy_balance')
    end subroutine

    !> surface_physics.f90:232-374
    subroutine mass_balance(now,par,bnd,day,year)
        type(surface_state_class), intent(in out) :: now
        type(boundary_opt_class),  intent(in)     :: bnd
        type(surface_param_class), intent(in)     :: par
        integer, intent(in) :: day, year
        type(surface_param_c) :: pc
        type(boundary_opt_c)  :: bc
        call par_to_c(par, pc); call bnd_to_c(bnd, bc)
        call check(semic_mass_balance_c(now%handle, pc, bc, int(day, c_int), int(year, c_int)), 'mass_balance')
    end subroutine

    !> surface_physics.f90:378-389, forwarded to the library (no host arithmetic here)
    subroutine sensible_heat_flux_v(ts, ta, wind, rhoa, csh, cap, shf)
        double precision, intent(in) :: ts(:), ta(:), wind(:), rhoa(:), csh, cap
        double precision, intent(out), target, contiguous :: shf(:)
        call check(semic_shf_c(ts, ta, wind, rhoa, csh, cap, c_loc(shf), int(size(ts), c_int)), 'sensible_heat_flux')
    end subroutine

    elemental subroutine sensible_heat_flux_e(ts, ta, wind, rhoa, csh, cap, shf)
        double precision, intent(in) :: ts, ta, wind, rhoa, csh, cap
        double precision, intent(out) :: shf
        shf = semic_shf_1(ts, ta, wind, rhoa, csh, cap)                    ! a pure procedure cannot stop: failure -> NaN
    end subroutine

    !> surface_physics.f90:393-430
    subroutine latent_heat_flux_v(ts, wind, shum, sp, rhoatm, mask, clh, eps, cls, clv, lhf, subl, evap)
        double precision, intent(in)  :: ts(:), wind(:), shum(:), sp(:), rhoatm(:), clh, eps, cls, clv
        integer,          intent(in)  :: mask(:)
        double precision, intent(out), target, contiguous :: lhf(:), subl(:), evap(:)
        call check(semic_lhf_c(ts, wind, shum, sp, rhoatm, int(mask, c_int), clh, eps, cls, clv, &
                               c_loc(lhf), c_loc(subl), c_loc(evap), int(size(ts), c_int)), 'latent_heat_flux')
    end subroutine

    elemental subroutine latent_heat_flux_e(ts, wind, shum, sp, rhoatm, mask, clh, eps, cls, clv, lhf, subl, evap)
        double precision, intent(in)  :: ts, shum, sp, rhoatm, wind, clh, eps, cls, clv
        integer,          intent(in)  :: mask
        double precision, intent(out) :: lhf, evap, subl
        lhf  = semic_lhf_1(ts, wind, shum, sp, rhoatm, int(mask, c_int), clh, eps, cls, clv, 0_c_int)
        subl = semic_lhf_1(ts, wind, shum, sp, rhoatm, int(mask, c_int), clh, eps, cls, clv, 1_c_int)
        evap = semic_lhf_1(ts, wind, shum, sp, rhoatm, int(mask, c_int), clh, eps, cls, clv, 2_c_int)
    end subroutine

    !> surface_physics.f90:434-438
    subroutine longwave_upward_v(ts, sigma, lwout)
        double precision, intent(in) :: ts(:), sigma
        double precision, intent(out), target, contiguous :: lwout(:)
        call check(semic_lwu_c(ts, sigma, c_loc(lwout), int(size(ts), c_int)), 'longwave_upward')
    end subroutine

    elemental subroutine longwave_upward_e(ts, sigma, lwout)
        double precision, intent(in) :: ts, sigma
        double precision, intent(out) :: lwout
        lwout = semic_lwu_1(ts, sigma)
    end subroutine

    !> surface_physics.f90:565-597 (device side; ave and now must both be allocated with surface_alloc)
    subroutine surface_physics_average(ave,now,step,nt)
        type(surface_state_class), intent(in out) :: ave
        type(surface_state_class), intent(in)     :: now
        character(len=*)  :: step
        double precision, optional, intent(in) :: nt
        real(c_double) :: ntc
        ntc = -1.0_c_double
        if (present(nt)) ntc = nt
        call check(semic_average_c(ave%handle, now%handle, trim(step)//c_null_char, ntc), 'surface_physics_average')
        if (trim(step) .eq. "end") &
            call check(semic_state_download(ave%handle, int(z'FFFFFFFF', c_int64_t)), 'surface_physics_average')
    end subroutine

    !> surface_physics.f90:633-677
    subroutine surface_alloc(now,npts)
        type(surface_state_class) :: now
        integer :: npts
        call check(semic_surface_alloc(now%handle, int(npts, c_int)), 'surface_alloc')
        call c_f_pointer(semic_state_field(now%handle, F_T2M), now%t2m, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_TSURF), now%tsurf, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_HSNOW), now%hsnow, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_HICE), now%hice, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_ALB), now%alb, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_ALB_SNOW), now%alb_snow, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_MELT), now%melt, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_MELTED_SNOW), now%melted_snow, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_MELTED_ICE), now%melted_ice, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_REFR), now%refr, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SMB), now%smb, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_ACC), now%acc, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_LHF), now%lhf, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SHF), now%shf, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_LWU), now%lwu, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SUBL), now%subl, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_EVAP), now%evap, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SMB_SNOW), now%smb_snow, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SMB_ICE), now%smb_ice, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_RUNOFF), now%runoff, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_QMR), now%qmr, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_QMR_RES), now%qmr_res, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_AMP), now%amp, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SF), now%sf, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_RF), now%rf, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SP), now%sp, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_LWD), now%lwd, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_SWD), now%swd, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_WIND), now%wind, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_RHOA), now%rhoa, [npts])
        call c_f_pointer(semic_state_field(now%handle, F_QQ), now%qq, [npts])
        call c_f_pointer(semic_state_mask(now%handle), now%mask, [npts])
    end subroutine

    !> surface_physics.f90:681-725
    subroutine surface_dealloc(now)
        type(surface_state_class), intent(in out) :: now
        call check(semic_surface_dealloc(now%handle), 'surface_dealloc')
        now%handle = c_null_ptr
        nullify(now%t2m, now%tsurf, now%hsnow, now%hice, now%alb, now%alb_snow, now%melt, now%melted_snow, &
                now%melted_ice, now%refr, now%smb, now%acc, now%lhf, now%shf, now%lwu, now%subl, now%evap, &
                now%smb_snow, now%smb_ice, now%runoff, now%qmr, now%qmr_res, now%amp, now%sf, now%rf, now%sp, &
                now%lwd, now%swd, now%wind, now%rhoa, now%qq, now%mask)
    end subroutine

    !> surface_physics.f90:729-814 (the C namelist reader keeps absent keys, sets tsticsub = tstic/n_ksub)
    subroutine surface_physics_par_load(par,filename)
        type(surface_param_class) :: par
        character(len=*)    :: filename
        type(surface_param_c) :: pc
        call par_to_c(par, pc)
        call check(semic_surface_physics_par_load(pc, trim(filename)//c_null_char), 'surface_physics_par_load')
        call par_from_c(pc, par)
    end subroutine

    !> surface_physics.f90:819-878
    subroutine surface_boundary_define(bnd,boundary)
        type(boundary_opt_class), intent(in out) :: bnd
        character(len=256), intent(in) :: boundary(:)
        type(boundary_opt_c) :: bc
        character(kind=c_char) :: buf(256*size(boundary))
        integer :: q, k
        do q = 1, size(boundary)
            do k = 1, 256
                buf((q-1)*256+k) = boundary(q)(k:k)
            end do
        end do
        call check(semic_surface_boundary_define(bc, buf, int(size(boundary), c_int), 256_c_int), 'surface_boundary_define')
        bnd%t2m = bc%t2m /= 0; bnd%tsurf = bc%tsurf /= 0; bnd%hsnow = bc%hsnow /= 0; bnd%alb = bc%alb /= 0
        bnd%melt = bc%melt /= 0; bnd%refr = bc%refr /= 0; bnd%smb = bc%smb /= 0; bnd%acc = bc%acc /= 0
        bnd%lhf = bc%lhf /= 0; bnd%shf = bc%shf /= 0; bnd%subl = bc%subl /= 0; bnd%amp = bc%amp /= 0
    end subroutine

    !> surface_physics.f90:882-915
    subroutine print_param(par)
        type(surface_param_class), intent(in) :: par
        type(surface_param_c) :: pc
        call par_to_c(par, pc)
        call check(semic_print_param_c(pc), 'print_param')
    end subroutine

    !> surface_physics.f90:919-932
    subroutine print_boundary_opt(bnd)
        type(boundary_opt_class), intent(in) :: bnd
        type(boundary_opt_c) :: bc
        call bnd_to_c(bnd, bc)
        call check(semic_print_boundary_opt_c(bc), 'print_boundary_opt')
    end subroutine

end module surface_physics
