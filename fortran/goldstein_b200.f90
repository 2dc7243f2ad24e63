! goldstein_b200.f90 -- MODULE goldstein_b200: step_goldstein with the reference's procedure name and argument list
! (src/goldstein/goldstein.f90:17-38), forwarding to the C-ABI.  The reference's MODULE goldstein stays in the build
! unchanged -- initialise_goldstein, end_goldstein, the restart and diagnostic routines are host code the coupler still
! calls (goldstein.f90:10-13) -- and only goldstein_wrapper's USE line changes (fortran/use_b200.py:
! `USE goldstein` -> `USE goldstein_b200, ONLY: step_goldstein`, genie_loop_wrappers.f90:122-123).  The step runs on the
! GPU; host arrays are touched only on output / restart steps (MOD(istep, iwstp|itstp|ianav|npstp) == 0).
MODULE goldstein_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  USE goldstein_lib, ONLY: maxi, maxj, maxk, maxl, npstp, iwstp, itstp, ianav
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: step_goldstein

CONTAINS

  SUBROUTINE step_goldstein(istep, latent_ocn, sensible_ocn, netsolar_ocn, &
       & netlong_ocn, fx0sic_ocn, evap_ocn, pptn_ocn, runoff_ocn, fwsic_ocn, &
       & stressxu_ocn, stressyu_ocn, stressxv_ocn, stressyv_ocn, tsval_ocn, &
       & ssval_ocn, usval_ocn, vsval_ocn, albedo_ocn, test_energy_ocean, &
       & test_water_ocean, go_ts, go_ts1, go_cost, go_u, go_tau, &
       & go_psi, go_mldta, go_rho)
    INTEGER, INTENT(IN) :: istep
    REAL, DIMENSION(:,:), INTENT(INOUT) :: &
         & latent_ocn, sensible_ocn, netsolar_ocn, netlong_ocn, fx0sic_ocn, &
         & evap_ocn, pptn_ocn, runoff_ocn, fwsic_ocn, stressxu_ocn, &
         & stressyu_ocn, stressxv_ocn, stressyv_ocn
    REAL, DIMENSION(:,:), INTENT(INOUT), TARGET, CONTIGUOUS :: tsval_ocn, ssval_ocn, usval_ocn, vsval_ocn, albedo_ocn
    REAL, INTENT(OUT), TARGET :: test_energy_ocean, test_water_ocean
    REAL, DIMENSION(:,:,:,:), INTENT(INOUT), TARGET, CONTIGUOUS :: go_ts, go_ts1
    REAL, DIMENSION(:,:), INTENT(INOUT), TARGET, CONTIGUOUS :: go_cost, go_mldta
    REAL, INTENT(INOUT), TARGET :: go_u(3,maxi,maxj,maxk)
    REAL, INTENT(INOUT) :: go_tau(2,maxi,maxj)
    REAL, INTENT(INOUT), TARGET :: go_psi(0:maxi,0:maxj)
    REAL, INTENT(INOUT), TARGET :: go_rho(maxi,maxj,maxk)

    TYPE(cg_goldstein_io), TARGET :: io
    LOGICAL :: output_step
    INTEGER(C_INT) :: rc

    CALL cg_ensure_handle()
    ! the surface fluxes and stresses of this step are already device resident: surflux, step_embm and
    ! step_seaice produced them on the GPU, so nothing is uploaded here.
    output_step = MOD(istep, npstp) == 0 .OR. MOD(istep, iwstp) == 0 .OR. &
         &        MOD(istep, itstp) == 0 .OR. MOD(istep, ianav) == 0
    IF (output_step) THEN
       io%tstar_ocn = C_LOC(tsval_ocn) ; io%sstar_ocn = C_LOC(ssval_ocn)
       io%ustar_ocn = C_LOC(usval_ocn) ; io%vstar_ocn = C_LOC(vsval_ocn)
       io%albedo_ocn = C_LOC(albedo_ocn)
       io%go_ts = C_NULL_PTR               ! download only: go_ts is not modified by the host between steps
       io%go_u = C_LOC(go_u) ; io%go_rho = C_LOC(go_rho) ; io%go_cost = C_NULL_PTR
       io%go_psi = C_LOC(go_psi)
       io%test_energy_ocean = C_LOC(test_energy_ocean) ; io%test_water_ocean = C_LOC(test_water_ocean)
       rc = cg_goldstein_step(cg_h, INT(istep, C_INT), C_LOC(io))
       CALL cg_check(rc, 'cg_goldstein_step')
       rc = cg_sync_to_host(cg_h, 'ts' // C_NULL_CHAR, 0_C_INT, C_LOC(go_ts), INT(SIZE(go_ts), C_INT64_T))
       CALL cg_check(rc, 'cg_sync_to_host(ts)')
       go_ts1 = go_ts
       rc = cg_goldstein_mldta(cg_h, 0_C_INT, C_LOC(go_mldta))     ! -5000 * mld; zero unless imld = 1
       CALL cg_check(rc, 'cg_goldstein_mldta')
    ELSE
       rc = cg_goldstein_step(cg_h, INT(istep, C_INT), C_NULL_PTR)
       CALL cg_check(rc, 'cg_goldstein_step')
    END IF
  END SUBROUTINE step_goldstein

END MODULE goldstein_b200