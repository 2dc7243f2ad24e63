! cgenie_b200_c.f90 -- ISO_C_BINDING interface to libcgenie_b200.so (include/cgenie_b200.h).
! In-tree precedent for BIND(C) interop in the reference: src/utils/itt_fortran.f90:7-41.
! Not compilable in the build container (no Fortran compiler there); built by users with gfortran:
!   python fortran/use_b200.py <cgenie>/src/wrappers/genie_loop_wrappers.f90        (switches the wrappers' USE lines)
!   gfortran -fdefault-real-8 -c cgenie_b200_c.f90 goldstein_b200.f90 embm_b200.f90 gold_seaice_b200.f90 biogem_b200.f90
!   ... link genie.exe with these objects next to the unchanged module objects and -L<repo>/cgenie_b200 -lcgenie_b200
MODULE cgenie_b200_c
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE

  TYPE(C_PTR), SAVE :: cg_h = C_NULL_PTR     ! one handle shared by the shim modules

  TYPE, BIND(C) :: cg_surflux_io
     TYPE(C_PTR) :: albedo_ocn, latent_ocn, sensible_ocn, netsolar_ocn, netlong_ocn, evap_ocn, precip_ocn, &
          & runoff_ocn, runoff_land, latent_atm, sensible_atm, netsolar_atm, netlong_atm, evap_atm, precip_atm, &
          & dhght_sic, dfrac_sic, temp_sic, albd_sic, qstar_atm
  END TYPE cg_surflux_io
  TYPE, BIND(C) :: cg_embm_io
     TYPE(C_PTR) :: tstar_atm, qstar_atm
  END TYPE cg_embm_io
  TYPE, BIND(C) :: cg_seaice_io
     TYPE(C_PTR) :: hght_sic, frac_sic, waterflux_ocn, conductflux_ocn
  END TYPE cg_seaice_io
  TYPE, BIND(C) :: cg_goldstein_io
     TYPE(C_PTR) :: tstar_ocn, sstar_ocn, ustar_ocn, vstar_ocn, albedo_ocn, go_ts, go_u, go_rho, go_cost, go_psi, &
          & test_energy_ocean, test_water_ocean
  END TYPE cg_goldstein_io

  INTERFACE
     INTEGER(C_INT) FUNCTION cg_create(jobdir, n_members, device, out) BIND(C, NAME='cg_create')
       IMPORT :: C_INT, C_CHAR, C_PTR
       CHARACTER(KIND=C_CHAR), DIMENSION(*), INTENT(IN) :: jobdir
       INTEGER(C_INT), VALUE :: n_members, device
       TYPE(C_PTR), INTENT(OUT) :: out
     END FUNCTION cg_create
     INTEGER(C_INT) FUNCTION cg_initialise(h) BIND(C, NAME='cg_initialise')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_initialise
     INTEGER(C_INT) FUNCTION cg_destroy(h) BIND(C, NAME='cg_destroy')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_destroy
     INTEGER(C_INT) FUNCTION cg_surflux_step(h, istep, io) BIND(C, NAME='cg_surflux_step')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h, io            ! io = C_LOC(a cg_surflux_io) or C_NULL_PTR: everything stays on the GPU
       INTEGER(C_INT), VALUE :: istep
     END FUNCTION cg_surflux_step
     INTEGER(C_INT) FUNCTION cg_embm_step(h, istep, io) BIND(C, NAME='cg_embm_step')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h, io
       INTEGER(C_INT), VALUE :: istep
     END FUNCTION cg_embm_step
     INTEGER(C_INT) FUNCTION cg_seaice_step(h, istep, io) BIND(C, NAME='cg_seaice_step')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h, io
       INTEGER(C_INT), VALUE :: istep
     END FUNCTION cg_seaice_step
     INTEGER(C_INT) FUNCTION cg_goldstein_step(h, istep, io) BIND(C, NAME='cg_goldstein_step')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h, io
       INTEGER(C_INT), VALUE :: istep
     END FUNCTION cg_goldstein_step
     INTEGER(C_INT) FUNCTION cg_goldstein_mldta(h, member, go_mldta) BIND(C, NAME='cg_goldstein_mldta')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h, go_mldta
       INTEGER(C_INT), VALUE :: member
     END FUNCTION cg_goldstein_mldta
     INTEGER(C_INT) FUNCTION cg_sync_to_host(h, name, member, dst, n) BIND(C, NAME='cg_sync_to_host')
       IMPORT :: C_INT, C_INT64_T, C_CHAR, C_PTR
       TYPE(C_PTR), VALUE :: h, dst
       CHARACTER(KIND=C_CHAR), DIMENSION(*), INTENT(IN) :: name
       INTEGER(C_INT), VALUE :: member
       INTEGER(C_INT64_T), VALUE :: n
     END FUNCTION cg_sync_to_host
     INTEGER(C_INT) FUNCTION cg_sync_from_host(h, name, member, src, n) BIND(C, NAME='cg_sync_from_host')
       IMPORT :: C_INT, C_INT64_T, C_CHAR, C_PTR
       TYPE(C_PTR), VALUE :: h, src
       CHARACTER(KIND=C_CHAR), DIMENSION(*), INTENT(IN) :: name
       INTEGER(C_INT), VALUE :: member
       INTEGER(C_INT64_T), VALUE :: n
     END FUNCTION cg_sync_from_host
     ! ---- BIOGEM / ATCHEM (include/cgenie_b200.h) ----
     INTEGER(C_INT) FUNCTION cg_biogem_forcing(h, genie_clock_ms) BIND(C, NAME='cg_biogem_forcing')
       IMPORT :: C_INT, C_INT64_T, C_PTR
       TYPE(C_PTR), VALUE :: h
       INTEGER(C_INT64_T), VALUE :: genie_clock_ms
     END FUNCTION cg_biogem_forcing
     INTEGER(C_INT) FUNCTION cg_biogem_step(h, dts, genie_clock_ms) BIND(C, NAME='cg_biogem_step')
       IMPORT :: C_INT, C_INT64_T, C_DOUBLE, C_PTR
       TYPE(C_PTR), VALUE :: h
       REAL(C_DOUBLE), VALUE :: dts
       INTEGER(C_INT64_T), VALUE :: genie_clock_ms
     END FUNCTION cg_biogem_step
     INTEGER(C_INT) FUNCTION cg_biogem_tracercoupling(h, go_ts, go_ts1) BIND(C, NAME='cg_biogem_tracercoupling')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h, go_ts, go_ts1      ! C_NULL_PTR = ts stays resident
     END FUNCTION cg_biogem_tracercoupling
     INTEGER(C_INT) FUNCTION cg_biogem_climate(h) BIND(C, NAME='cg_biogem_climate')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_biogem_climate
     INTEGER(C_INT) FUNCTION cg_biogem_climate_sol(h) BIND(C, NAME='cg_biogem_climate_sol')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_biogem_climate_sol
     INTEGER(C_INT) FUNCTION cg_atchem_step(h, dts) BIND(C, NAME='cg_atchem_step')
       IMPORT :: C_INT, C_DOUBLE, C_PTR
       TYPE(C_PTR), VALUE :: h
       REAL(C_DOUBLE), VALUE :: dts
     END FUNCTION cg_atchem_step
     INTEGER(C_INT) FUNCTION cg_cpl_flux_ocnatm(h) BIND(C, NAME='cg_cpl_flux_ocnatm')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_cpl_flux_ocnatm
     INTEGER(C_INT) FUNCTION cg_cpl_flux_ocnsed(h, dts) BIND(C, NAME='cg_cpl_flux_ocnsed')
       IMPORT :: C_INT, C_PTR, C_DOUBLE
       TYPE(C_PTR), VALUE :: h
       REAL(C_DOUBLE), VALUE :: dts
     END FUNCTION cg_cpl_flux_ocnsed
     INTEGER(C_INT) FUNCTION cg_cpl_comp_ocnsed(h, ocnstep, mbiogem, msedgem) BIND(C, NAME='cg_cpl_comp_ocnsed')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
       INTEGER(C_INT), VALUE :: ocnstep, mbiogem, msedgem
     END FUNCTION cg_cpl_comp_ocnsed
     INTEGER(C_INT) FUNCTION cg_reinit_flux_rokocn(h) BIND(C, NAME='cg_reinit_flux_rokocn')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_reinit_flux_rokocn
     INTEGER(C_INT) FUNCTION cg_biogem_sig_update(h, dts, ben_dmin) BIND(C, NAME='cg_biogem_sig_update')
       IMPORT :: C_INT, C_PTR, C_DOUBLE
       TYPE(C_PTR), VALUE :: h
       REAL(C_DOUBLE), VALUE :: dts, ben_dmin
     END FUNCTION cg_biogem_sig_update
     ! the export / air-sea flux / "misc" integrals of diag_biogem_timeseries as well (field "bg_sig2"; biogem.f90:2870-2883, 2926-2964,
     ! 3058-3062): call once, before the first BIOGEM step whose window integrals are wanted (initialise_biogem's place)
     INTEGER(C_INT) FUNCTION cg_biogem_sig_extended(h) BIND(C, NAME='cg_biogem_sig_extended')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_biogem_sig_extended
     INTEGER(C_INT) FUNCTION cg_biogem_sig_reset(h) BIND(C, NAME='cg_biogem_sig_reset')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_biogem_sig_reset
     ! diag_biogem_timeslice's arithmetic (biogem.f90:2421-2699): 3-D carbonate re-solve + window integrals on the device
     INTEGER(C_INT) FUNCTION cg_biogem_slice_update(h, dts) BIND(C, NAME='cg_biogem_slice_update')
       IMPORT :: C_INT, C_PTR, C_DOUBLE
       TYPE(C_PTR), VALUE :: h
       REAL(C_DOUBLE), VALUE :: dts
     END FUNCTION cg_biogem_slice_update
     INTEGER(C_INT) FUNCTION cg_biogem_slice_reset(h) BIND(C, NAME='cg_biogem_slice_reset')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: h
     END FUNCTION cg_biogem_slice_reset
  END INTERFACE

CONTAINS

  ! Non-zero status -> the reference's own failure path (src/wrappers/genie_util.f90:15-33)
  SUBROUTINE cg_check(rc, where)
    USE genie_global, ONLY: write_status
    INTEGER(C_INT), INTENT(IN) :: rc
    CHARACTER(LEN=*), INTENT(IN) :: where
    IF (rc /= 0) THEN
       PRINT *, 'cgenie_b200 error ', rc, ' in ', where
       CALL write_status('ERRORED')
    END IF
  END SUBROUTINE cg_check

  ! Lazily create the device model from the job directory genie.exe runs in ('.')
  SUBROUTINE cg_ensure_handle()
    INTEGER(C_INT) :: rc
    INTEGER :: n_members
    CHARACTER(LEN=32) :: env
    IF (.NOT. C_ASSOCIATED(cg_h)) THEN
       n_members = 1
       CALL GET_ENVIRONMENT_VARIABLE('CGENIE_B200_MEMBERS', env)
       IF (LEN_TRIM(env) > 0) READ (env, *) n_members
       rc = cg_create('.' // C_NULL_CHAR, INT(n_members, C_INT), 0_C_INT, cg_h)
       CALL cg_check(rc, 'cg_create')
       rc = cg_initialise(cg_h)
       CALL cg_check(rc, 'cg_initialise')
    END IF
  END SUBROUTINE cg_ensure_handle

END MODULE cgenie_b200_c
