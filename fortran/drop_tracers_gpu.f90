! drop_tracers_gpu.f90 -- how the reference's drop_tracers.f90 calls the C ABI (see equilibration_gpu.f90).
! update_tracer_population + init (drop_tracers.f90:29-35) become lbg_mp_init, the propagate loop
! (:41-49) becomes lbg_mp_step.  Not compiled in this repository's image (no Fortran compiler).
subroutine drop_tracers_gpu(h)
  use, intrinsic :: iso_c_binding
  use precision_kinds, only: dp
  use module_input, only: getinput
  use laboetie_gpu
  implicit none
  type(c_ptr), intent(in) :: h
  integer(c_int) :: rc, done, conv
  integer :: it, maxsteps, i, chunk
  real(c_double) :: vacf0(3), f_ext(3)
  real(c_double), allocatable :: vacf(:, :)

  maxsteps = getinput%int("maximum_moment_propagation_steps", 0)
  if (maxsteps == 0) return                                          ! drop_tracers.f90:20-21
  if (maxsteps < 0) maxsteps = huge(1)                               ! :40
  f_ext = getinput%dp3('f_ext', [0._dp, 0._dp, 0._dp])               ! :92
  rc = lbg_mp_init(h, getinput%dp('tracer_Db', 0._dp), getinput%dp('tracer_ka', 0._dp), &
                   getinput%dp('tracer_kd', 0._dp), f_ext, vacf0)
  if (rc == LBG_ERR_TRACER_DB) error stop 'The diffusion coefficient (tracer_Db in input file) is invalid'
  if (rc /= LBG_OK) error stop "lbg_mp_init"
  open (99, file='output/vacf.dat')
  write (99, *) '# time t, VACF_x(t), VACF_y(t), VACF_z(t)'
  write (99, *) 0, vacf0                                             ! module_moment_propagation.f90:145
  chunk = 4096
  allocate (vacf(3, chunk))
  it = 0
  conv = 0
  do while (it < maxsteps .and. conv == 0)
    rc = lbg_mp_step(h, min(chunk, maxsteps - it), vacf, done, conv)
    if (rc == LBG_ERR_RESTPART_NEGATIVE) stop 'somewhere restpart is negative'
    if (rc /= LBG_OK) error stop "lbg_mp_step"
    do i = 1, done
      write (99, *) it + i, vacf(:, i)                               ! module_moment_propagation.f90:269
    end do
    it = it + done
  end do
  close (99)
end subroutine drop_tracers_gpu
