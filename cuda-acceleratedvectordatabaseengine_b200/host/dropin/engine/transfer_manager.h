// Drop-in seat of the reference's engine/transfer_manager.h (vdb::TransferManager lives in the mirror header).
#pragma once
#include "../../ivf_flat_index.h"
