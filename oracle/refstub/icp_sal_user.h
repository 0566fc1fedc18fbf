/* refstub/icp_sal_user.h — fake QAT driver (see cpa.h). TEST INFRASTRUCTURE ONLY. */
#ifndef REFSTUB_ICP_SAL_USER_H
#define REFSTUB_ICP_SAL_USER_H
#include "cpa.h"
CpaStatus icp_sal_userStart(const char *pProcessName);
CpaStatus icp_sal_userStop(void);
CpaBoolean icp_sal_userIsQatAvailable(void);
#endif
