! biogem_b200.f90 -- MODULE biogem_b200 / atchem_b200 / sedgem_b200 / rokgem_b200: the loop entry points of BIOGEM, ATCHEM and
! the SEDGEM / ROKGEM coupler calls with the reference's procedure names and argument lists (src/biogem/biogem.f90:528-547,
! 1885-1890, 2083-2087, 2132-2150, 2243-2247; src/atchem/atchem.f90:63-67, 252-282, 306-320; src/sedgem/sedgem.f90:894-937,
! 1029-1068; src/rokgem/rokgem.f90:472-480), forwarding to the C-ABI.  The reference's own modules stay in the build
! unchanged (MODULE biogem exports 15 names -- initialise_biogem, diag_biogem_timeslice, biogem_save_restart, end_biogem ...,
! biogem.f90:6-21 -- that the ini / end / diagnostic wrappers keep calling); fortran/use_b200.py switches the USE lines of the
! wrappers listed there (genie_loop_wrappers.f90:178-183, 197-203, 219-226, 289-293, 310-342, 439-468).
! The biogeochemistry runs on the GPU.  The padded interface arrays the coupler owns are filled from the device's compact
! fields after every BIOGEM step (dum_sfcocn1 <- "sfcocn1", dum_sfxsed1 <- "sfxsed1": one member, 36 x 36 x (16 + 9) doubles)
! so that unchanged Fortran consumers (diag_biogem_timeslice / _timeseries, SEDGEM) read what the reference would give them;
! b200_fill_interface = .FALSE. restricts this to the caller's save steps (the copies join the device streams).
! Not compiled here: no Fortran compiler exists in the build image (DESIGN.md 1).
MODULE biogem_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  USE gem_cmn, ONLY: n_l_ocn, n_l_sed, conv_iselected_io, conv_iselected_is
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: step_biogem, biogem_tracercoupling, biogem_forcing, biogem_climate, biogem_climate_sol
  LOGICAL, PUBLIC, SAVE :: b200_fill_interface = .TRUE.

CONTAINS

  SUBROUTINE biogem_forcing(dum_genie_clock)
    INTEGER(KIND=8), INTENT(IN) :: dum_genie_clock
    CALL cg_ensure_handle()
    CALL cg_check(cg_biogem_forcing(cg_h, INT(dum_genie_clock, C_INT64_T)), 'cg_biogem_forcing')
  END SUBROUTINE biogem_forcing

  SUBROUTINE step_biogem(dum_dts, dum_genie_clock, dum_sfcatm1, dum_sfxatm1, &
       & dum_sfcocn1, dum_sfxocn1, dum_sfcsed1, dum_sfxsed1, dum_sfxsumrok1)
    REAL, INTENT(IN) :: dum_dts
    INTEGER(KIND=8), INTENT(IN) :: dum_genie_clock
    REAL, INTENT(IN),    DIMENSION(:,:,:) :: dum_sfcatm1
    REAL, INTENT(INOUT), DIMENSION(:,:,:) :: dum_sfxatm1
    REAL, INTENT(OUT),   DIMENSION(:,:,:) :: dum_sfcocn1
    REAL, INTENT(INOUT), DIMENSION(:,:,:) :: dum_sfxocn1
    REAL, INTENT(IN),    DIMENSION(:,:,:) :: dum_sfcsed1
    REAL, INTENT(INOUT), DIMENSION(:,:,:) :: dum_sfxsed1
    REAL, INTENT(INOUT), DIMENSION(:,:,:) :: dum_sfxsumrok1
    REAL(C_DOUBLE), ALLOCATABLE, TARGET :: buf(:,:,:)
    INTEGER :: l
    CALL cg_ensure_handle()
    CALL cg_check(cg_biogem_step(cg_h, REAL(dum_dts, C_DOUBLE), INT(dum_genie_clock, C_INT64_T)), 'cg_biogem_step')
    ! the air-sea flux is accumulated into sfxsumatm on the device (cpl_flux_ocnatm, atchem.f90:306-320, is fused into the
    ! kernel): the host copy handed on to cpl_flux_ocnatm_wrapper carries no flux
    dum_sfxatm1 = 0.0
    IF (b200_fill_interface) THEN
       ! bottom-water composition and sediment rain of this step (biogem.f90:1724-1761), compact device order -> padded arrays
       ALLOCATE(buf(n_l_ocn, SIZE(dum_sfcocn1, 2), SIZE(dum_sfcocn1, 3)))     ! field "sfcocn1": Fortran shape (n_l_ocn, maxi, maxj)
       CALL cg_check(cg_sync_to_host(cg_h, 'sfcocn1' // C_NULL_CHAR, 0_C_INT, C_LOC(buf), INT(SIZE(buf), C_INT64_T)), 'cg_sync_to_host(sfcocn1)')
       DO l = 1, n_l_ocn
          dum_sfcocn1(conv_iselected_io(l),:,:) = buf(l,:,:)
       END DO
       DEALLOCATE(buf)
       ALLOCATE(buf(n_l_sed, SIZE(dum_sfxsed1, 2), SIZE(dum_sfxsed1, 3)))
       CALL cg_check(cg_sync_to_host(cg_h, 'sfxsed1' // C_NULL_CHAR, 0_C_INT, C_LOC(buf), INT(SIZE(buf), C_INT64_T)), 'cg_sync_to_host(sfxsed1)')
       DO l = 1, n_l_sed
          dum_sfxsed1(conv_iselected_is(l),:,:) = buf(l,:,:)
       END DO
       DEALLOCATE(buf)
    END IF
  END SUBROUTINE step_biogem

  SUBROUTINE biogem_tracercoupling(dum_ts, dum_ts1)
    REAL, DIMENSION(:,:,:,:), INTENT(INOUT) :: dum_ts, dum_ts1
    CALL cg_ensure_handle()
    CALL cg_check(cg_biogem_tracercoupling(cg_h, C_NULL_PTR, C_NULL_PTR), 'cg_biogem_tracercoupling')
  END SUBROUTINE biogem_tracercoupling

  SUBROUTINE biogem_climate(dum_hght_sic, dum_frac_sic, dum_cost, dum_solfor, dum_fxsw, dum_uvw, dum_tau, dum_psi, &
       & dum_uv, dum_usurf, dum_mld, dum_evap, dum_pptn, dum_solconst)
    REAL, DIMENSION(:,:), INTENT(IN) :: dum_hght_sic, dum_frac_sic
    REAL, DIMENSION(:,:), INTENT(INOUT) :: dum_cost
    REAL, DIMENSION(:), INTENT(IN) :: dum_solfor
    REAL, DIMENSION(:,:), INTENT(IN) :: dum_fxsw
    REAL, DIMENSION(:,:,:,:), INTENT(IN) :: dum_uvw
    REAL, DIMENSION(:,:,:), INTENT(IN) :: dum_tau
    REAL, DIMENSION(:,:), INTENT(IN) :: dum_psi
    REAL, DIMENSION(:,:,:), INTENT(IN) :: dum_uv
    REAL, DIMENSION(:,:), INTENT(IN) :: dum_usurf, dum_mld, dum_evap, dum_pptn
    REAL, INTENT(INOUT) :: dum_solconst
    CALL cg_ensure_handle()
    CALL cg_check(cg_biogem_climate(cg_h), 'cg_biogem_climate')   ! physics is aliased on the device; resets go_cost there
    dum_cost = 0.0
  END SUBROUTINE biogem_climate

  SUBROUTINE biogem_climate_sol(dum_solfor, dum_fxsw, dum_solconst)
    REAL, DIMENSION(:), INTENT(IN) :: dum_solfor
    REAL, DIMENSION(:,:), INTENT(IN) :: dum_fxsw
    REAL, INTENT(INOUT) :: dum_solconst
    CALL cg_ensure_handle()
    CALL cg_check(cg_biogem_climate_sol(cg_h), 'cg_biogem_climate_sol')
  END SUBROUTINE biogem_climate_sol

END MODULE biogem_b200
MODULE atchem_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: step_atchem, cpl_flux_ocnatm, cpl_comp_atmocn, cpl_comp_EMBM

CONTAINS

  SUBROUTINE step_atchem(dum_dts, dum_sfxsumatm, dum_sfcatm)
    REAL, INTENT(IN) :: dum_dts
    REAL, DIMENSION(:,:,:), INTENT(INOUT) :: dum_sfxsumatm
    REAL, DIMENSION(:,:,:), INTENT(OUT), TARGET, CONTIGUOUS :: dum_sfcatm
    CALL cg_ensure_handle()
    CALL cg_check(cg_atchem_step(cg_h, REAL(dum_dts, C_DOUBLE)), 'cg_atchem_step')   ! includes cpl_comp_atmocn
    dum_sfxsumatm = 0.0
  END SUBROUTINE step_atchem

  SUBROUTINE cpl_flux_ocnatm(dum_dts, dum_sfxatm1, dum_sfxsumatm)
    REAL, INTENT(IN) :: dum_dts
    REAL, DIMENSION(:,:,:), INTENT(INOUT) :: dum_sfxatm1, dum_sfxsumatm
    CALL cg_check(cg_cpl_flux_ocnatm(cg_h), 'cg_cpl_flux_ocnatm')   ! already applied by cg_biogem_step
    dum_sfxatm1 = 0.0
  END SUBROUTINE cpl_flux_ocnatm

  SUBROUTINE cpl_comp_atmocn(dum_n_atm, dum_sfcatm, dum_sfcatm1)
    INTEGER, INTENT(IN) :: dum_n_atm
    REAL, DIMENSION(:,:,:), INTENT(IN) :: dum_sfcatm
    REAL, DIMENSION(:,:,:), INTENT(INOUT) :: dum_sfcatm1
    ! fused into cg_atchem_step: the ocean-grid copy "sfcatm1" is device resident
  END SUBROUTINE cpl_comp_atmocn

  SUBROUTINE cpl_comp_EMBM(dum_t, dum_q, dum_sfcatm1)
    REAL, DIMENSION(:,:), INTENT(IN) :: dum_t, dum_q
    REAL, DIMENSION(:,:,:), INTENT(INOUT) :: dum_sfcatm1
    ! rows 1-2 of the device's "sfcatm1" are filled behind the ATCHEM step (k_bg_cpl_comp_embm); the host copy follows the
    ! reference (atchem.f90:270-282) so that host-side readers see the same values
    dum_sfcatm1(1,:,:) = dum_t
    dum_sfcatm1(2,:,:) = dum_q
  END SUBROUTINE cpl_comp_EMBM

END MODULE atchem_b200
! SEDGEM / ROKGEM coupler routines genie.f90 calls after every BIOGEM step whether or not the modules run
! (genie.f90:413-427 through genie_loop_wrappers.f90:197-226, 289-293).  The interface arrays are device resident
! (fields "sfxsumsed", "sfcsumocn", "sfxsumrok1"); the host arrays are left alone.  Jobs that really run SEDGEM or ROKGEM
! keep the reference's modules and exchange the arrays through cg_sync_to_host / cg_sync_from_host.
MODULE sedgem_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: cpl_flux_ocnsed, cpl_comp_ocnsed

CONTAINS

  SUBROUTINE cpl_flux_ocnsed(dum_dts, dum_n_maxsed, dum_n_maxi, dum_n_maxj, dum_ns_maxi, dum_ns_maxj, dum_sfxsed1, dum_sfxsumsed)
    REAL, INTENT(IN) :: dum_dts
    INTEGER, INTENT(IN) :: dum_n_maxsed, dum_n_maxi, dum_n_maxj, dum_ns_maxi, dum_ns_maxj
    REAL, DIMENSION(dum_n_maxsed,dum_n_maxi,dum_n_maxj), INTENT(INOUT) :: dum_sfxsed1
    REAL, DIMENSION(dum_n_maxsed,dum_ns_maxi,dum_ns_maxj), INTENT(INOUT) :: dum_sfxsumsed
    IF (dum_ns_maxi /= dum_n_maxi .OR. dum_ns_maxj /= dum_n_maxj) CALL cg_check(2_C_INT, 'cpl_flux_ocnsed: sediment grid /= ocean grid')
    CALL cg_check(cg_cpl_flux_ocnsed(cg_h, REAL(dum_dts, C_DOUBLE)), 'cg_cpl_flux_ocnsed')
  END SUBROUTINE cpl_flux_ocnsed

  SUBROUTINE cpl_comp_ocnsed(dum_ocnstep, dum_mbiogem, dum_msedgem, dum_n_i_ocn, dum_n_j_ocn, dum_n_i_sed, dum_n_j_sed, &
       & dum_sfcocn1, dum_sfcsumocn)
    INTEGER, INTENT(IN) :: dum_ocnstep, dum_mbiogem, dum_msedgem, dum_n_i_ocn, dum_n_j_ocn, dum_n_i_sed, dum_n_j_sed
    REAL, DIMENSION(:,:,:), INTENT(IN) :: dum_sfcocn1
    REAL, DIMENSION(:,:,:), INTENT(INOUT) :: dum_sfcsumocn
    IF (dum_n_i_sed /= dum_n_i_ocn .OR. dum_n_j_sed /= dum_n_j_ocn) CALL cg_check(2_C_INT, 'cpl_comp_ocnsed: sediment grid /= ocean grid')
    CALL cg_check(cg_cpl_comp_ocnsed(cg_h, INT(dum_ocnstep, C_INT), INT(dum_mbiogem, C_INT), INT(dum_msedgem, C_INT)), 'cg_cpl_comp_ocnsed')
  END SUBROUTINE cpl_comp_ocnsed

END MODULE sedgem_b200
MODULE rokgem_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: reinit_flux_rokocn

CONTAINS

  SUBROUTINE reinit_flux_rokocn(dum_sfxsumrok1)
    REAL, DIMENSION(:,:,:), INTENT(INOUT) :: dum_sfxsumrok1
    dum_sfxsumrok1 = 0.0
    CALL cg_check(cg_reinit_flux_rokocn(cg_h), 'cg_reinit_flux_rokocn')
  END SUBROUTINE reinit_flux_rokocn

END MODULE rokgem_b200