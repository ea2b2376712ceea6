! laboetie_gpu_iface.f90 -- ISO_C_BINDING interfaces to include/laboetie_gpu.h.
!
! The reference driver (main.f90, init_simu.f90, equilibration.f90, drop_tracers.f90) stays Fortran;
! add this module to the Makefile list (Makefile:18-47), link liblaboetie_gpu.so, and replace the loop
! bodies as shown in fortran/equilibration_gpu.f90 and fortran/drop_tracers_gpu.f90.
!
! NOTE: there is no Fortran compiler in the image this repository is built and tested in, so this
! file has not been compiled here.  The same symbols are exercised through the C++ mirror of the
! driver (laboetie_b200/driver) and the ctypes binding (laboetie_b200/api.py).
module laboetie_gpu
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: lbg_create, lbg_create_slab, lbg_create_geometry, lbg_get_nature, lbg_destroy, lbg_comm_unique_id, lbg_comm_init, lbg_partition
  public :: lbg_get_interfacial, lbg_get_counts
  public :: lbg_lb_set_in_place, lbg_lb_init, lbg_lb_upload, lbg_lb_set_force_uniform, lbg_lb_set_force_field, lbg_lb_step, lbg_lb_time
  public :: lbg_lb_download_moments, lbg_lb_download_populations, lbg_lb_profiles, lbg_lb_total_flux, lbg_lb_probe
  public :: lbg_lb_download_moments_async, lbg_wait_transfers, lbg_get_info, lbg_lb_slice
  public :: lbg_mp_init, lbg_mp_init_from_moments, lbg_mp_step, lbg_mp_download, lbg_sync, lbg_status_message

  integer(c_int), parameter, public :: LBG_OK = 0
  integer(c_int), parameter, public :: LBG_ERR_NEGATIVE_POPULATION = 1   ! equilibration.f90:248
  integer(c_int), parameter, public :: LBG_ERR_RESTPART_NEGATIVE = 2     ! module_moment_propagation.f90:257
  integer(c_int), parameter, public :: LBG_ERR_RELAXATION_TIME = 3       ! module_collision.f90:39
  integer(c_int), parameter, public :: LBG_ERR_TRACER_DB = 4             ! drop_tracers.f90:89

  interface
    integer(c_int) function lbg_create(h, lx, ly, lz, nature, device) bind(C, name="lbg_create")
      import :: c_ptr, c_int, c_int8_t
      type(c_ptr), intent(out) :: h
      integer(c_int), value :: lx, ly, lz, device
      integer(c_int8_t), intent(in) :: nature(*)          ! node%nature copied to a contiguous (lx,ly,lz) array
    end function
    integer(c_int) function lbg_create_slab(h, lx, ly, lz_global, k0, nzl, nature_halo, device) bind(C, name="lbg_create_slab")
      import :: c_ptr, c_int, c_int8_t
      type(c_ptr), intent(out) :: h
      integer(c_int), value :: lx, ly, lz_global, k0, nzl, device
      integer(c_int8_t), intent(in) :: nature_halo(*)
    end function
    integer(c_int) function lbg_create_geometry(h, label, lx, ly, lz_global, k0, nzl, device) bind(C, name="lbg_create_geometry")
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: h
      integer(c_int), value :: label, lx, ly, lz_global, k0, nzl, device
    end function
    integer(c_int) function lbg_get_nature(h, nature) bind(C, name="lbg_get_nature")
      import :: c_ptr, c_int, c_int8_t
      type(c_ptr), value :: h
      integer(c_int8_t), intent(out) :: nature(*)
    end function
    integer(c_int) function lbg_destroy(h) bind(C, name="lbg_destroy")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
    end function
    integer(c_int) function lbg_comm_unique_id(id) bind(C, name="lbg_comm_unique_id")
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
    end function
    integer(c_int) function lbg_comm_init(h, nranks, rank, id) bind(C, name="lbg_comm_init")
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: h
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id(128)
    end function
    integer(c_int) function lbg_partition(lz, nranks, rank, k0, nzl) bind(C, name="lbg_partition")
      import :: c_int
      integer(c_int), value :: lz, nranks, rank
      integer(c_int), intent(out) :: k0, nzl
    end function
    integer(c_int) function lbg_get_interfacial(h, interfacial) bind(C, name="lbg_get_interfacial")
      import :: c_ptr, c_int, c_int8_t
      type(c_ptr), value :: h
      integer(c_int8_t), intent(out) :: interfacial(*)    ! logical is 4 bytes in Fortran: convert with /= 0
    end function
    integer(c_int) function lbg_get_counts(h, n_fluid, n_interfacial_fluid) bind(C, name="lbg_get_counts")
      import :: c_ptr, c_int, c_int64_t
      type(c_ptr), value :: h
      integer(c_int64_t), intent(out) :: n_fluid, n_interfacial_fluid
    end function
    integer(c_int) function lbg_lb_set_in_place(h, on) bind(C, name="lbg_lb_set_in_place")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int), value :: on
    end function
    integer(c_int) function lbg_lb_init(h, rho0) bind(C, name="lbg_lb_init")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), value :: rho0
    end function
    integer(c_int) function lbg_lb_upload(h, n, rho, jx, jy, jz) bind(C, name="lbg_lb_upload")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: n(*), rho(*), jx(*), jy(*), jz(*)   ! n(i,j,k,l) exactly as system::n
    end function
    integer(c_int) function lbg_lb_set_force_uniform(h, f) bind(C, name="lbg_lb_set_force_uniform")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: f(3)
    end function
    integer(c_int) function lbg_lb_set_force_field(h, fx, fy, fz) bind(C, name="lbg_lb_set_force_field")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: fx(*), fy(*), fz(*)
    end function
    integer(c_int) function lbg_lb_step(h, tau, nsteps, check_every, target_error, l2err_hist, steps_done, converged) &
        bind(C, name="lbg_lb_step")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), value :: tau, target_error
      integer(c_int), value :: nsteps, check_every
      real(c_double), intent(out) :: l2err_hist(*)
      integer(c_int), intent(out) :: steps_done, converged
    end function
    integer(c_int) function lbg_lb_time(h, t) bind(C, name="lbg_lb_time")
      import :: c_ptr, c_int, c_int64_t
      type(c_ptr), value :: h
      integer(c_int64_t), intent(out) :: t
    end function
    integer(c_int) function lbg_lb_download_moments(h, rho, jx, jy, jz) bind(C, name="lbg_lb_download_moments")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: rho(*), jx(*), jy(*), jz(*)
    end function
    ! the same read-back, queued: rho, jx, jy, jz are valid after lbg_wait_transfers (Phase B may run meanwhile)
    integer(c_int) function lbg_lb_download_moments_async(h, rho, jx, jy, jz) bind(C, name="lbg_lb_download_moments_async")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: rho(*), jx(*), jy(*), jz(*)
    end function
    integer(c_int) function lbg_wait_transfers(h) bind(C, name="lbg_wait_transfers")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
    end function
    integer(c_int) function lbg_get_info(h, key, val) bind(C, name="lbg_get_info")
      import :: c_ptr, c_int, c_char, c_int64_t
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: key(*)       ! NUL-terminated, e.g. "mp_neighbour_table"//c_null_char
      integer(c_int64_t), intent(out) :: val
    end function
    integer(c_int) function lbg_lb_download_populations(h, n) bind(C, name="lbg_lb_download_populations")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: n(*)
    end function
    integer(c_int) function lbg_lb_profiles(h, axis, raw, out) bind(C, name="lbg_lb_profiles")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: axis, raw
      real(c_double), intent(out) :: out(*)
    end function
    integer(c_int) function lbg_lb_total_flux(h, out) bind(C, name="lbg_lb_total_flux")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: out(3)
    end function
    ! one plane of density / momentum density (equilibration.f90:526-548); axis 0: x = index -> (ly, nzl), ...
    integer(c_int) function lbg_lb_slice(h, axis, index, rho, jx, jy, jz) bind(C, name="lbg_lb_slice")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: axis, index
      real(c_double), intent(out) :: rho(*), jx(*), jy(*), jz(*)
    end function
    integer(c_int) function lbg_lb_probe(h, i, j, k, out) bind(C, name="lbg_lb_probe")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: i, j, k                      ! 0-based
      real(c_double), intent(out) :: out(4)
    end function
    integer(c_int) function lbg_mp_init(h, Db, ka, kd, f_ext, vacf0) bind(C, name="lbg_mp_init")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), value :: Db, ka, kd
      real(c_double), intent(in) :: f_ext(3)
      real(c_double), intent(out) :: vacf0(3)
    end function
    ! Phase B from the driver's own arrays (node%solventdensity, node%solventflux of drop_tracers.f90:63-105)
    integer(c_int) function lbg_mp_init_from_moments(h, rho, jx, jy, jz, Db, ka, kd, f_ext, vacf0) &
        bind(C, name="lbg_mp_init_from_moments")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: rho(*), jx(*), jy(*), jz(*)
      real(c_double), value :: Db, ka, kd
      real(c_double), intent(in) :: f_ext(3)
      real(c_double), intent(out) :: vacf0(3)
    end function
    integer(c_int) function lbg_mp_step(h, nsteps, vacf, steps_done, converged) bind(C, name="lbg_mp_step")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: nsteps
      real(c_double), intent(out) :: vacf(3, *)
      integer(c_int), intent(out) :: steps_done, converged
    end function
    integer(c_int) function lbg_mp_download(h, P, Pads) bind(C, name="lbg_mp_download")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: P(3, *), Pads(3, *)    ! Propagated_Quantity(x:z,i,j,k,now)
    end function
    integer(c_int) function lbg_sync(h) bind(C, name="lbg_sync")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
    end function
    type(c_ptr) function lbg_status_string_c(status) bind(C, name="lbg_status_string")
      import :: c_ptr, c_int
      integer(c_int), value :: status
    end function
  end interface

contains

  ! the reference's own stop messages, keyed by status
  function lbg_status_message(status) result(msg)
    integer(c_int), intent(in) :: status
    character(len=:), allocatable :: msg
    character(kind=c_char), pointer :: p(:)
    integer :: n
    call c_f_pointer(lbg_status_string_c(status), p, [256])
    n = 0
    do while (p(n + 1) /= c_null_char .and. n < 255)
      n = n + 1
    end do
    allocate (character(len=n) :: msg)
    msg = transfer(p(1:n), msg)
  end function

end module laboetie_gpu
