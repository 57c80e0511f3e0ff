// global.cc -- see global.H; defaults of src/global.cc:63-74
#include "global.H"

namespace exphost {

double dtime = 0.01, tnow = 0.0;
int    this_step = 0;
unsigned multistep = 0;
int    Mstep = 1, mstep = 0, mdrft = 0;
std::vector<int> mintvl, mfirst;
std::vector<std::vector<bool>> mactive;
std::vector<std::vector<int>>  dstepL, dstepN;
double dynfracS = 1.0, dynfracD = 1.0e32, dynfracV = 0.01, dynfracA = 0.03, dynfracP = 0.05;
int    shiftlevl = 0, centerlevl = -1;
bool   DTold = false, initializing = false, use_cuda = true;
int    myid = 0, numprocs = 1, cudaDevice = 0;
void  *nccl_comm = nullptr;
ComponentContainer *comp = nullptr;

void reset_globals()
{
  dtime = 0.01; tnow = 0.0; this_step = 0; multistep = 0; Mstep = 1; mstep = 0; mdrft = 0;
  mintvl.clear(); mfirst.clear(); mactive.clear(); dstepL.clear(); dstepN.clear();
  dynfracS = 1.0; dynfracD = 1.0e32; dynfracV = 0.01; dynfracA = 0.03; dynfracP = 0.05;
  shiftlevl = 0; centerlevl = -1; DTold = false; initializing = false; use_cuda = true;
}

} // namespace exphost
