! Drop-in replacement for xtt-lib-fortran/elliptic_tools.f90, PROMOTED build (the whole driver compiled with -freal-4-real-8:
! every real(4) is a real(8), so the module binds the _f64 entry points).
!
! Same module name, same public procedures, same argument order, kinds and intents
! (reference elliptic_tools.f90:8-14, 64-69, 93-109, 333-335), so src/diagnose/main.f90 and
! every include compile unchanged.  The bodies forward through ISO_C_BINDING to the
! C-ABI library of this repository (include/xee_b200.h, libxee_b200.so), which runs
! the work in hand-written CUDA on the GPU.  Build: fortran/build_fortran.sh.
!
! The abandoned solve_elliptic_AC (reference :267-331, never called, divides by
! uninitialised values) is intentionally not provided.
module elliptic_tools
use iso_c_binding
implicit none
integer, parameter :: err_over_max_iteration = ISHFT(1,0), &
    &                 err_explode = ISHFT(1,1)

interface
    subroutine xee_cal_coe_f64(a, b, c, coe, dx, dy, nx, ny, err) bind(C, name="xee_cal_coe_f64")
        import :: c_double, c_int
        real(c_double), intent(in)    :: a(*), b(*), c(*), dx, dy
        real(c_double), intent(inout) :: coe(*)
        integer(c_int), intent(in)   :: nx, ny
        integer(c_int), intent(inout):: err
    end subroutine
    subroutine xee_do_elliptic_f64(psi, coe, outdat, nx, ny, err) bind(C, name="xee_do_elliptic_f64")
        import :: c_double, c_int
        real(c_double), intent(in)    :: psi(*), coe(*)
        real(c_double), intent(inout) :: outdat(*)
        integer(c_int), intent(in)   :: nx, ny
        integer(c_int), intent(inout):: err
    end subroutine
    subroutine xee_solve_elliptic_f64(max_iter, check_step, converge_time, lost_rate, strategy_r1, strategy_r2, &
    &                                 alpha, dat, coe, f, workspace, nx, ny, err, debug) &
    &                                 bind(C, name="xee_solve_elliptic_f64")
        import :: c_double, c_int
        integer(c_int), intent(inout):: max_iter, err
        integer(c_int), intent(in)   :: check_step, converge_time, lost_rate, nx, ny, debug
        real(c_double), intent(inout) :: strategy_r1, strategy_r2, dat(*), workspace(*)
        real(c_double), intent(in)    :: alpha, coe(*), f(*)
    end subroutine
    subroutine xee_judge_error(err) bind(C, name="xee_judge_error")
        import :: c_int
        integer(c_int), intent(in) :: err
    end subroutine
end interface

contains

subroutine cal_coe(a, b, c, workspace, dx, dy, nx, ny, err)
implicit none
real(8), intent(in) :: a(nx-1, ny-2), b(nx-1, ny-1), c(nx-2, ny-1), &
&                      dx, dy
real(8), intent(inout) :: workspace(9, nx, ny)
integer, intent(in)    :: nx, ny
integer, intent(inout) :: err
call xee_cal_coe_f64(a, b, c, workspace, dx, dy, nx, ny, err)
end subroutine

subroutine do_elliptic(psi, coe, outdat, nx, ny, err)
implicit none
real(8), intent(in)    :: psi(nx, ny), coe(9, nx,ny)
real(8), intent(inout) :: outdat(nx,ny)
integer, intent(in)    :: nx, ny
integer, intent(inout) :: err
call xee_do_elliptic_f64(psi, coe, outdat, nx, ny, err)
end subroutine

subroutine solve_elliptic(max_iter, check_step, converge_time, lost_rate, &
&                         strategy_r1, strategy_r2, alpha, dat, coe, f, &
&                         workspace, nx, ny, err, debug)
implicit none
real(8), intent(inout), target :: dat(nx,ny), workspace(nx, ny)
real(8), intent(inout) :: strategy_r1, strategy_r2
real(8), intent(in)    :: coe(9, nx, ny), f(nx, ny), alpha
integer, intent(in)    :: nx, ny, check_step, converge_time, lost_rate, debug
integer, intent(inout) :: max_iter, err
call xee_solve_elliptic_f64(max_iter, check_step, converge_time, lost_rate, strategy_r1, strategy_r2, &
&                           alpha, dat, coe, f, workspace, nx, ny, err, debug)
end subroutine

subroutine judge_error(err)
implicit none
integer, intent(in) :: err
call xee_judge_error(err)
end subroutine

end module elliptic_tools
