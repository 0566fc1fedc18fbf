/* refstub/icp_sal_poll.h — fake QAT driver (see cpa.h). TEST INFRASTRUCTURE ONLY. */
#ifndef REFSTUB_ICP_SAL_POLL_H
#define REFSTUB_ICP_SAL_POLL_H
#include "cpa.h"
CpaStatus icp_sal_DcPollInstance(CpaInstanceHandle instanceHandle, Cpa32U responseQuota);
#endif
