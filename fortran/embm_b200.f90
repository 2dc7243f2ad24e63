! embm_b200.f90 -- MODULE embm_b200: step_embm and surflux with the reference's names and argument lists
! (src/embm/embm.f90:22-38, 2548-2588).  MODULE embm itself stays in the build (initialise_embm, end_embm, radfor ...,
! embm.f90:13-19); fortran/use_b200.py switches the USE lines of surflux_wrapper and embm_wrapper
! (genie_loop_wrappers.f90:7-8, 61-62).  Only the arrays the coupler reads on the host are filled, and only on output steps.
MODULE embm_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  USE embm_lib, ONLY: maxi, maxj, npstp, iwstp, itstp, ianav, ndta
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: step_embm, surflux

CONTAINS

  SUBROUTINE step_embm(istep, latent_atm, sensible_atm, netsolar_atm, &
       & netlong_atm, evap_atm, pptn_atm, stressxu_atm, stressyu_atm, &
       & stressxv_atm, stressyv_atm, tstar_atm, qstar_atm, &
       & torog_atm, surf_orog_atm, flag_ents, lowestlu2_atm, lowestlv3_atm)
    INTEGER, INTENT(IN) :: istep
    REAL, DIMENSION(:,:), INTENT(IN) :: latent_atm, sensible_atm, netsolar_atm, netlong_atm, pptn_atm, evap_atm
    REAL, DIMENSION(:,:), INTENT(OUT) :: stressxu_atm, stressyu_atm, stressxv_atm, stressyv_atm
    REAL, DIMENSION(:,:), INTENT(OUT), TARGET, CONTIGUOUS :: tstar_atm, qstar_atm
    REAL, INTENT(OUT) :: torog_atm(maxi,maxj)
    REAL, INTENT(IN) :: surf_orog_atm(maxi,maxj)
    LOGICAL, INTENT(IN) :: flag_ents
    REAL, DIMENSION(:,:), INTENT(INOUT) :: lowestlu2_atm, lowestlv3_atm
    TYPE(cg_embm_io), TARGET :: io
    INTEGER(C_INT) :: rc
    CALL cg_ensure_handle()
    IF (MOD(istep, itstp * ndta) == 0 .OR. MOD(istep, iwstp * ndta) == 0 .OR. MOD(istep, npstp * ndta) == 0) THEN
       io%tstar_atm = C_LOC(tstar_atm) ; io%qstar_atm = C_LOC(qstar_atm)
       rc = cg_embm_step(cg_h, INT(istep, C_INT), C_LOC(io))
       torog_atm = tstar_atm
    ELSE
       rc = cg_embm_step(cg_h, INT(istep, C_INT), C_NULL_PTR)
    END IF
    CALL cg_check(rc, 'cg_embm_step')
  END SUBROUTINE step_embm

  SUBROUTINE surflux(istep, otemp, osaln, atemp, ashum, sich, sica, &
       & tice, albice, stressxu_ocn, stressyu_ocn, &
       & stressxv_ocn, stressyv_ocn, albedo, fxlho, fxsho, fxswo, fxlwo, &
       & evap_ocn, pptn_ocn, runoff_ocn, runoff_land, fxlha, fxsha, &
       & fxswa, fxlwa, evap_atm, pptn_atm, dthsic, dtareasic, &
       & atmos_lowestlh_atm, go_solfor, go_fxsw, dum_sfcatm, &
       & eb_ca, gn_daysperyear, eb_fx0a, eb_fx0o, eb_fxsen, eb_fxlw, &
       & eb_evap, eb_pptn, eb_relh, eb_uv, eb_usurf, solconst, &
       & co2_out, ch4_out, n2o_out, surf_orog_atm, landice_slicemask_lic, &
       & albs_atm, land_albs_snow_lnd, land_albs_nosnow_lnd, &
       & land_snow_lnd, land_bcap_lnd, land_z0_lnd, land_temp_lnd, &
       & land_moisture_lnd, flag_ents, lowestlu2_atm, lowestlv3_atm)
    INTEGER, INTENT(IN) :: istep
    REAL, DIMENSION(:,:), INTENT(IN) :: otemp, osaln, atemp, sich, sica, &
         & stressxu_ocn, stressyu_ocn, stressxv_ocn, stressyv_ocn
    REAL, DIMENSION(:,:), INTENT(OUT) :: ashum, tice, albice, &
         & albedo, fxlho, fxsho, fxswo, fxlwo, pptn_ocn, &
         & runoff_ocn, runoff_land, fxlha, fxsha, fxswa, fxlwa, &
         & dthsic, dtareasic, co2_out, ch4_out, n2o_out, &
         & evap_ocn, evap_atm, pptn_atm, atmos_lowestlh_atm
    REAL, INTENT(OUT) :: go_solfor(maxj), go_fxsw(maxi,maxj)
    REAL, INTENT(IN), DIMENSION(:,:,:) :: dum_sfcatm
    REAL, INTENT(IN) :: gn_daysperyear
    REAL, DIMENSION(:,:), INTENT(IN) :: eb_ca
    REAL, DIMENSION(:,:), INTENT(OUT) :: eb_fx0a, eb_fx0o, eb_fxsen, eb_fxlw, eb_evap, eb_pptn, eb_relh, eb_usurf
    REAL, INTENT(OUT) :: eb_uv(2,maxi,maxj)
    REAL, INTENT(IN) :: solconst
    REAL, INTENT(INOUT) :: surf_orog_atm(maxi,maxj)
    REAL, DIMENSION(:,:), INTENT(OUT) :: landice_slicemask_lic
    REAL, DIMENSION(:,:), INTENT(INOUT) :: albs_atm, land_snow_lnd, land_bcap_lnd, land_z0_lnd, &
         & land_temp_lnd, land_moisture_lnd
    REAL, DIMENSION(:,:), INTENT(IN) :: land_albs_snow_lnd, land_albs_nosnow_lnd
    LOGICAL, INTENT(IN) :: flag_ents
    REAL, DIMENSION(:,:), INTENT(INOUT) :: lowestlu2_atm, lowestlv3_atm
    INTEGER(C_INT) :: rc
    CALL cg_ensure_handle()
    ! all ~40 flux fields stay on the GPU: their consumers (step_embm, step_seaice, step_goldstein) are
    ! device resident too.  Diagnostics that want them call cg_sync_to_host(name) at output intervals.
    rc = cg_surflux_step(cg_h, INT(istep, C_INT), C_NULL_PTR)
    CALL cg_check(rc, 'cg_surflux_step')
  END SUBROUTINE surflux

END MODULE embm_b200