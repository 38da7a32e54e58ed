! ref_replay -- replays a fixture written by export_fixture.py through the REFERENCE's own routing routines
! (IRF_route_basin, accum_inst_runoff, irf_rch, kwt_rch and everything they call), compiled from /root/reference in place.
! Written without a Fortran compiler at hand: see README.md.  usage: ref_replay <fixture.txt> <out.txt>
program ref_replay
  use nrtype
  use public_var
  use dataTypes,           only: RCHTOPO, RCHPRP, STRFLX, STRSTA, dlength
  use globalData,          only: FRAC_FUTURE, nRoutes, routeMethods, onRoute, idxSUM, idxIRF, idxKWT
  use process_param,       only: basinUH, make_uh
  use basinUH_module,      only: IRF_route_basin
  use accum_runoff_module, only: accum_runoff_rch
  use irf_route_module,    only: irf_route_rch
  use kwt_route_module,    only: kwt_route_rch
  implicit none
  character(len=1024)               :: fixture, outFile
  character(len=strLen)             :: cmessage
  integer(i4b)                      :: nRch, nHRU, nSteps, ierr, i, k, t, iRoute, nUps, ntdh, gb
  integer(i4b)                      :: hwDrain
  real(dp)                          :: fshape, tscale, velo, diff, mannN, wScale, minLen, T0, T1
  type(RCHTOPO), allocatable        :: NETOPO(:)
  type(RCHPRP),  allocatable        :: RPARAM(:)
  type(STRFLX),  allocatable        :: RCHFLX(:)
  type(STRSTA),  allocatable        :: RCHSTA(:)
  type(dlength), allocatable        :: segUH(:)
  real(dp),      allocatable        :: rlen(:), qi(:)
  type(accum_runoff_rch)            :: routeSUM
  type(irf_route_rch)               :: routeIRF
  type(kwt_route_rch)               :: routeKWT

  call get_command_argument(1, fixture)
  call get_command_argument(2, outFile)
  open(11, file=trim(fixture), status='old', action='read')
  open(12, file=trim(outFile), status='replace', action='write')

  read(11,*) nRch, nHRU, nSteps, nRoutes
  read(11,*) dt, fshape, tscale, velo, diff, mannN, wScale, hwDrain, minLen
  allocate(routeMethods(nRoutes))
  read(11,*) routeMethods(1:nRoutes)

  ! options the routines read from public_var (everything else keeps its declared default)
  hw_drain_point   = hwDrain
  min_length_route = minLen
  is_lake_sim      = .false.
  is_flux_wm       = .false.
  qmodOption       = 0
  ! method slots, as read_control.f90:583-597
  onRoute(:) = .false.
  idxSUM = -1; idxIRF = -1; idxKWT = -1
  do iRoute = 1, nRoutes
    select case (routeMethods(iRoute))
      case (accumRunoff);           idxSUM = iRoute; onRoute(accumRunoff) = .true.
      case (impulseResponseFunc);   idxIRF = iRoute; onRoute(impulseResponseFunc) = .true.
      case (kinematicWaveTracking); idxKWT = iRoute; onRoute(kinematicWaveTracking) = .true.
      case default
        write(*,*) 'ref_replay: only route_opt 0 / 1 / 2 are replayed'; stop 2
    end select
  end do

  ! ---- network in processing order with the products of augment_ntopo the routines read
  allocate(NETOPO(nRch), RPARAM(nRch), RCHFLX(nRch), RCHSTA(nRch), rlen(nRch), qi(nRch))
  do i = 1, nRch
    read(11,*) NETOPO(i)%REACHID, NETOPO(i)%DREACHI, RPARAM(i)%RLENGTH, RPARAM(i)%R_SLOPE, RPARAM(i)%BASAREA, RPARAM(i)%UPSAREA, &
               RPARAM(i)%TOTAREA, RPARAM(i)%R_WIDTH, nUps
    NETOPO(i)%REACHIX = i
    NETOPO(i)%RHORDER = i
    NETOPO(i)%DREACHK = -1
    if (NETOPO(i)%DREACHI > 0) then
      continue                                  ! DREACHK is filled below, once every id has been read
    else
      NETOPO(i)%DREACHI = -1                    ! outlet (network_topo.f90: no downstream reach)
    end if
    NETOPO(i)%LAKINLT = .false.; NETOPO(i)%ISLAKE = .false.; NETOPO(i)%LAKETARGVOL = .false.; NETOPO(i)%LAKEMODELTYPE = 0
    RPARAM(i)%R_MAN_N = mannN
    RPARAM(i)%MINFLOW = 0._dp
    rlen(i) = RPARAM(i)%RLENGTH
    allocate(NETOPO(i)%UREACHI(nUps), NETOPO(i)%UREACHK(nUps), NETOPO(i)%goodBas(nUps))
    do k = 1, nUps
      read(11,*) NETOPO(i)%UREACHI(k), gb
      NETOPO(i)%goodBas(k) = (gb == 1)
    end do
  end do
  do i = 1, nRch
    if (NETOPO(i)%DREACHI > 0) NETOPO(i)%DREACHK = NETOPO(NETOPO(i)%DREACHI)%REACHID
    do k = 1, size(NETOPO(i)%UREACHI)
      NETOPO(i)%UREACHK(k) = NETOPO(NETOPO(i)%UREACHI(k))%REACHID
    end do
  end do

  ! ---- unit hydrographs by the reference's own routines (process_param.f90)
  call basinUH(dt, fshape, tscale, ierr, cmessage); call check(ierr, cmessage)
  if (onRoute(impulseResponseFunc)) then
    call make_uh(rlen, dt, velo, diff, segUH, ierr, cmessage); call check(ierr, cmessage)
    do i = 1, nRch
      allocate(NETOPO(i)%UH(size(segUH(i)%dat)))
      NETOPO(i)%UH(:) = segUH(i)%dat(:)
    end do
  end if

  ! ---- cold start, init_model_data.f90:398-463
  do i = 1, nRch
    allocate(RCHFLX(i)%ROUTE(nRoutes))
    RCHFLX(i)%BASIN_QI = 0._dp; RCHFLX(i)%BASIN_QR(0:1) = 0._dp
    RCHFLX(i)%Qelapsed = 0; RCHFLX(i)%Qobs = 0._dp
    RCHFLX(i)%REACH_WM_FLUX = 0._dp; RCHFLX(i)%REACH_WM_VOL = 0._dp      ! main_route.f90:110-123 (water management off)
    RCHFLX(i)%basinEvapo = 0._dp; RCHFLX(i)%basinPrecip = 0._dp
    do iRoute = 1, nRoutes
      RCHFLX(i)%ROUTE(iRoute)%REACH_VOL(0:1) = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%REACH_Q        = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%REACH_INFLOW   = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%REACH_WM_FLUX_actual = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%WB             = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%Qerror         = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%FLOOD_VOL(0:1) = 0._dp
      RCHFLX(i)%ROUTE(iRoute)%REACH_ELE      = 0._dp
    end do
    if (onRoute(impulseResponseFunc)) then
      ntdh = size(NETOPO(i)%UH)
      allocate(RCHFLX(i)%QFUTURE_IRF(ntdh))
      RCHFLX(i)%QFUTURE_IRF(:) = 0._dp
    end if
  end do

  ! ---- the steps: main_route.f90:205-266 with doesBasinRoute = 1, one thread, all reaches
  T0 = 0._dp; T1 = dt                                 ! init_model_data.f90:600
  do t = 1, nSteps
    read(11,*) qi(1:nRch)
    do i = 1, nRch
      RCHFLX(i)%BASIN_QI = qi(i)
    end do
    call IRF_route_basin(NETOPO, RCHFLX, ierr, cmessage); call check(ierr, cmessage)
    do iRoute = 1, nRoutes
      do i = 1, nRch                                  ! processing order: upstream reaches first
        select case (routeMethods(iRoute))
          case (accumRunoff)
            call routeSUM%route(i, T0, T1, NETOPO, RPARAM, RCHSTA, RCHFLX, ierr, cmessage)
          case (impulseResponseFunc)
            call routeIRF%route(i, T0, T1, NETOPO, RPARAM, RCHSTA, RCHFLX, ierr, cmessage)
          case (kinematicWaveTracking)
            call routeKWT%route(i, T0, T1, NETOPO, RPARAM, RCHSTA, RCHFLX, ierr, cmessage)
        end select
        call check(ierr, cmessage)
      end do
      write(12,'(*(ES25.17E3,1X))') (RCHFLX(i)%ROUTE(iRoute)%REACH_Q, i = 1, nRch)
    end do
    T0 = T1; T1 = T0 + dt                             ! init_model_data.f90:311-312
  end do
  close(11); close(12)

contains
  subroutine check(err, msg)
    integer(i4b), intent(in) :: err
    character(*), intent(in) :: msg
    if (err /= 0) then
      write(*,'(A,I0,1X,A)') 'ref_replay: ierr ', err, trim(msg)
      stop 1
    end if
  end subroutine check
end program ref_replay
