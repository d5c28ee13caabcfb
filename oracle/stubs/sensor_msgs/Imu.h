// TEST INFRASTRUCTURE stub (type not used by the code that is compiled)
#pragma once
