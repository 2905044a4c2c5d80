#pragma once
namespace Sophus { class SO3 {}; class SE3 {}; }
