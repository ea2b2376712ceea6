! equilibration_gpu.f90 -- how the reference's equilibration.f90 calls the C ABI.
!
! This is the patch a laboetie maintainer applies, shown as a stand-alone subroutine.  Everything the
! reference does outside the loop body is kept with its cadence: the step print (equilibration.f90:149),
! the profile dumps at t == 1 and every print_files_frequency steps from the density / momentum the step
! starts from (:154-176), v_centralnode.dat (:185-188), total_mass_flux.dat (:259-261), l2err.dat (:344), the
! two-stage convergence state machine with the uniform force (:377-386) or the compensated particle force
! (:388-487), the final profiles and fields (:493-548) and the write-back to node%... (:551-555).  The body of
! `do t=1,HUGE(t)` (:194-343: collide, bounce-back, streaming, ANY(n<0), density, momentum, l2err) becomes one
! call that runs up to the next step at which the driver has something to write.
! laboetie_b200/driver/laboetie_driver.cpp is the compiled C++ mirror of this control flow (tests/test_driver.py);
! this file itself is not compiled in this repository's image (no Fortran compiler).
subroutine equilibration_gpu(h)
  use, intrinsic :: iso_c_binding
  use precision_kinds, only: dp
  use system, only: fluid, node, supercell
  use module_input, only: getinput
  use constants, only: x, y, z
  use laboetie_gpu
  implicit none
  type(c_ptr), intent(out) :: h            ! kept by the caller: drop_tracers_gpu continues from the resident state
  integer(c_int) :: rc, done, conv
  integer(c_int8_t), allocatable :: nature(:, :, :)
  real(c_double), allocatable :: density(:, :, :), jx(:, :, :), jy(:, :, :), jz(:, :, :), hist(:)
  real(c_double), allocatable :: f_ext_x(:, :, :), f_ext_y(:, :, :), f_ext_z(:, :, :)
  real(c_double) :: f_ext_loc(3), tau, target_error, probe(4), flux(3)
  real(dp), parameter :: eps = epsilon(1._dp)
  integer :: n1, n2, n3, t, tfext, i, j, k, chunk, next_dump, print_frequency, print_files_frequency
  integer :: px, py, pz
  logical :: convergence_reached_without_fext, compensate_f_ext, write_total_mass_flux

  n1 = getinput%int("lx", assert=">0"); n2 = getinput%int("ly", assert=">0"); n3 = getinput%int("lz", assert=">0")
  tau = getinput%dp('relaxation_time', defaultvalue=1._dp, assert=">0")
  target_error = getinput%dp("target_error", 1.D-10)
  print_frequency = getinput%int('print_frequency', defaultvalue=max(int(50000/(n1*n2*n3)), 1), assert=">0")
  print_files_frequency = getinput%int("print_files_frequency", HUGE(1))
  compensate_f_ext = getinput%log("compensate_f_ext", .false.)
  write_total_mass_flux = getinput%log("write_total_mass_flux", .false.)
  allocate (nature(n1, n2, n3), source=int(node%nature, c_int8_t))       ! equilibration.f90:93
  allocate (density(n1, n2, n3), jx(n1, n2, n3), jy(n1, n2, n3), jz(n1, n2, n3))

  rc = lbg_create(h, n1, n2, n3, nature, 0_c_int);                 if (rc /= LBG_OK) error stop "lbg_create"
  rc = lbg_lb_init(h, getinput%dp("initialSolventDensity", 1._dp)) ! init_simu.f90:24-39
  open (13, file="./output/l2err.dat")
  if (compensate_f_ext) open (79, file="./output/v_centralnode.dat")
  if (write_total_mass_flux) open (65, file="output/total_mass_flux.dat")
  ! units 56-58 / 66-68 (density and mass-flux profiles) are opened as in equilibration.f90:127-138

  allocate (hist(4096))
  convergence_reached_without_fext = .false.
  px = 0; py = 0; pz = 0; tfext = 0
  t = 0
  do
    ! ---- what the reference does at the top of step t+1 (equilibration.f90:149-188) -------------------------
    if (modulo(t + 1, print_files_frequency) == 0 .or. t + 1 == 1) call dump_profiles(t + 1)
    if (compensate_f_ext .and. convergence_reached_without_fext) then            ! :185-188
      rc = lbg_lb_probe(h, int(px - 1, c_int), int(py - 1, c_int), int(pz - 1, c_int), probe)
      write (79, *) t + 1 - tfext, probe(1), probe(2), probe(3)
    end if
    if (write_total_mass_flux) then                                               ! :259-261 (momentum of step t)
      rc = lbg_lb_total_flux(h, flux)
      write (65, *) t + 1, real(flux)
    end if
    ! ---- steps t+1 .. t+chunk: never past the step before the next thing to write ----------------------------
    chunk = 4096
    next_dump = (t/print_files_frequency + 1)*print_files_frequency              ! smallest multiple > t
    if (next_dump - 1 - t >= 1) chunk = min(chunk, next_dump - 1 - t)
    if (t + 1 == next_dump) chunk = min(chunk, print_files_frequency)
    if (write_total_mass_flux .or. (compensate_f_ext .and. convergence_reached_without_fext)) chunk = 1
    rc = lbg_lb_step(h, tau, int(chunk, c_int), 1_c_int, target_error, hist, done, conv)
    do i = 1, done
      write (13, *) t + i, hist(i)                                                ! :344
      if (modulo(t + i, print_frequency) == 0) print *, t + i, real(hist(i)), "(target", real(target_error, 4), ")"
    end do
    t = t + done
    if (rc == LBG_ERR_NEGATIVE_POPULATION) error stop "In equilibration, the population n(x,y,z,vel) < 0"
    if (rc /= LBG_OK) error stop "lbg_lb_step"
    if (conv == 0) cycle
    if (convergence_reached_without_fext) exit                                    ! :373-374
    convergence_reached_without_fext = .true.                                     ! :377-386
    tfext = t + 1
    f_ext_loc = getinput%dp3("f_ext", [0._dp, 0._dp, 0._dp])
    if (.not. compensate_f_ext) then
      rc = lbg_lb_set_force_uniform(h, f_ext_loc)
    else
      ! equilibration.f90:388-487 stays host code exactly as it is in the reference: the patch only moves that block
      ! (particle diameter / parity checks, particle_coordinates, the loop that marks the particle nodes, the
      ! geometryLabel == -1 background compensation, the zeroing on solid nodes) into the contained procedure
      ! build_compensated_force below, which fills the three per-node arrays and returns the particle centre.
      allocate (f_ext_x(n1, n2, n3), f_ext_y(n1, n2, n3), f_ext_z(n1, n2, n3))
      call build_compensated_force(f_ext_loc, f_ext_x, f_ext_y, f_ext_z, px, py, pz)
      rc = lbg_lb_set_force_field(h, f_ext_x, f_ext_y, f_ext_z)
    end if
  end do
  close (13)
  if (compensate_f_ext) close (79)
  if (write_total_mass_flux) close (65)

  call dump_profiles(-1)                                                          ! :493-521, "# Steady state ..."
  rc = lbg_lb_download_moments(h, density, jx, jy, jz)                             ! :551-554
  ! (a driver that goes straight on to drop_tracers can queue this read-back with lbg_lb_download_moments_async and
  !  call lbg_wait_transfers before it first touches the arrays: Phase B works on the resident copies)
  open (69, file="output/mass-flux_field_2d_at_x.eq.1.dat")                      ! :526-532
  do j = 1, n2
    do k = 1, n3
      write (69, *) j, k, jy(1, j, k), jz(1, j, k)
    end do
  end do
  close (69)
  node%solventdensity = density
  node%solventflux(x) = jx
  node%solventflux(y) = jy
  node%solventflux(z) = jz

contains

  ! body = equilibration.f90:388-487, unchanged (not repeated in this repository)
  subroutine build_compensated_force(f_loc, fx, fy, fz, cx, cy, cz)
    real(c_double), intent(in) :: f_loc(3)
    real(c_double), intent(out) :: fx(:, :, :), fy(:, :, :), fz(:, :, :)
    integer, intent(out) :: cx, cy, cz
    ! ... the reference's block, with f_ext_x/y/z -> fx/fy/fz and px/py/pz -> cx/cy/cz ...
  end subroutine build_compensated_force

  ! equilibration.f90:154-176 (step >= 1) and :493-521 (step < 0): plane sums on the device, one call per axis
  subroutine dump_profiles(step)
    integer, intent(in) :: step
    real(c_double), allocatable :: prof(:, :)
    integer :: axis, u, p, np
    do axis = 2, 0, -1                           ! z -> units 66/56, y -> 67/57, x -> 68/58
      u = 2 - axis
      np = merge(n3, merge(n2, n1, axis == 1), axis == 2)
      allocate (prof(4, np))
      rc = lbg_lb_profiles(h, int(axis, c_int), 0_c_int, prof)
      if (step >= 1) then
        write (66 + u, *) "# timestep", step
        write (56 + u, *) "# timestep", step
      else
        write (66 + u, *) "# Steady state with convergence criteria", real(target_error)
        write (56 + u, *) "# Steady state with convergence criteria", real(target_error)
      end if
      do p = 1, np
        write (66 + u, *) p, prof(1, p), prof(2, p), prof(3, p)
        write (56 + u, *) p, prof(4, p)
      end do
      if (step >= 1) write (66 + u, *)
      deallocate (prof)
    end do
  end subroutine dump_profiles
end subroutine equilibration_gpu
