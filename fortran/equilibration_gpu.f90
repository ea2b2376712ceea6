! equilibration_gpu.f90 -- how the reference's equilibration.f90 calls the C ABI.
!
! This is the patch a laboetie maintainer applies, shown as a stand-alone subroutine.  Everything the
! reference does outside the loop body (file opens, print_frequency, profile dumps, the two-stage
! convergence state machine, the write-back to node%...) is kept; the body of `do t=1,HUGE(t)`
! (equilibration.f90:194-343: collide, bounce-back, streaming, ANY(n<0), density, momentum, l2err)
! becomes one call.  Not compiled in this repository's image (no Fortran compiler).
subroutine equilibration_gpu
  use, intrinsic :: iso_c_binding
  use precision_kinds, only: dp
  use system, only: fluid, node, supercell
  use module_input, only: getinput
  use constants, only: x, y, z
  use laboetie_gpu
  implicit none
  type(c_ptr) :: h
  integer(c_int) :: rc, done, conv
  integer(c_int8_t), allocatable :: nature(:, :, :)
  real(c_double), allocatable :: density(:, :, :), jx(:, :, :), jy(:, :, :), jz(:, :, :), hist(:), prof(:, :)
  real(c_double) :: f_ext_loc(3), tau, target_error
  integer :: n1, n2, n3, t, i, k, chunk
  logical :: convergence_reached_without_fext

  n1 = getinput%int("lx", assert=">0"); n2 = getinput%int("ly", assert=">0"); n3 = getinput%int("lz", assert=">0")
  tau = getinput%dp('relaxation_time', defaultvalue=1._dp, assert=">0")
  target_error = getinput%dp("target_error", 1.D-10)
  allocate (nature(n1, n2, n3), source=int(node%nature, c_int8_t))       ! equilibration.f90:93
  allocate (density(n1, n2, n3), jx(n1, n2, n3), jy(n1, n2, n3), jz(n1, n2, n3))

  rc = lbg_create(h, n1, n2, n3, nature, 0_c_int);                 if (rc /= LBG_OK) error stop "lbg_create"
  rc = lbg_lb_init(h, getinput%dp("initialSolventDensity", 1._dp)) ! init_simu.f90:24-39
  open (13, file="./output/l2err.dat")

  chunk = 4096
  allocate (hist(chunk), prof(4, n3))
  convergence_reached_without_fext = .false.
  t = 0
  do
    ! equilibration.f90:154-176 -- profile dumps use the density / momentum the step starts from
    if (t == 0) then
      rc = lbg_lb_profiles(h, 2_c_int, 0_c_int, prof)
      do k = 1, n3
        write (66, *) k, prof(1, k), prof(2, k), prof(3, k)
        write (56, *) k, prof(4, k)
      end do
    end if
    ! equilibration.f90:194-343 for up to `chunk` steps; returns at the first converged step
    rc = lbg_lb_step(h, tau, chunk, 1_c_int, target_error, hist, done, conv)
    do i = 1, done
      write (13, *) t + i, hist(i)                                  ! equilibration.f90:344
    end do
    t = t + done
    if (rc == LBG_ERR_NEGATIVE_POPULATION) error stop "In equilibration, the population n(x,y,z,vel) < 0"
    if (rc /= LBG_OK) error stop "lbg_lb_step"
    if (conv == 0) cycle
    if (.not. convergence_reached_without_fext) then               ! equilibration.f90:377-386
      convergence_reached_without_fext = .true.
      f_ext_loc = getinput%dp3("f_ext", [0._dp, 0._dp, 0._dp])
      rc = lbg_lb_set_force_uniform(h, f_ext_loc)
    else
      exit                                                          ! equilibration.f90:373-374
    end if
  end do

  rc = lbg_lb_download_moments(h, density, jx, jy, jz)              ! equilibration.f90:551-554
  node%solventdensity = density
  node%solventflux(x) = jx
  node%solventflux(y) = jy
  node%solventflux(z) = jz
  ! keep `h` (module variable in the real patch): drop_tracers_gpu continues from the resident state
end subroutine equilibration_gpu
