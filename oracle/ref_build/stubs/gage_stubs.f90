! Stand-ins for gageMeta_data / obs_data (which read csv / netCDF files): globalData declares one variable of each type.
module gageMeta_data
  implicit none
  type, public :: gageMeta
    integer :: dummy = 0
  end type gageMeta
end module gageMeta_data

module obs_data
  implicit none
  type, public :: gageObs
    integer :: dummy = 0
  end type gageObs
  type, public :: waterTake
    integer :: dummy = 0
  end type waterTake
end module obs_data
